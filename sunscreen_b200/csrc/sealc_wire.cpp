// sealc_wire.cpp — the reference's wire format (Save / SaveSize / Load / UnsafeLoad) for the data objects that cross
// the seal_fhe FFI: Ciphertext, Plaintext, PublicKey, SecretKey, KSwitchKeys (RelinKeys / GaloisKeys).
//
// Format (S/serialization.h:96-113, S/serialization.cpp:236-446): every object is
//     SEALHeader { u16 magic 0xA15E; u8 header_size 0x10; u8 major; u8 minor; u8 compr_mode; u16 reserved; u64 size }
// followed by its members; with compr_mode zlib (1) / zstd (2) the members are compressed as ONE stream and `size`
// counts header + compressed bytes.  Nested objects (the DynArray inside a ciphertext, the ciphertexts inside key
// lists, the PRNG info of a seeded ciphertext) are always written with their own header and compr_mode none.
//   Ciphertext members   S/ciphertext.cpp:186-238   parms_id[4], u8 is_ntt, u64 size, u64 n, u64 k, f64 scale,
//                                                   u64 correction_factor, DynArray<u64> data  (seeded: first
//                                                   polynomial only + UniformRandomGeneratorInfo, :204-224)
//   Plaintext members    S/plaintext.cpp:204-229    parms_id[4], u64 coeff_count, f64 scale, DynArray<u64>
//   DynArray members     S/dynarray.h:652-680       u64 size, size * sizeof(T) bytes
//   KSwitchKeys members  S/kswitchkeys.cpp:42-84    parms_id[4], u64 dim1, per list: u64 dim2, dim2 x Ciphertext
//   PRNG info members    S/randomgen.cpp:99-121     u8 prng_type, 64-byte seed
//   PublicKey = its Ciphertext (S/publickey.h:89-110), SecretKey = its Plaintext (S/secretkey.h).
// Loading validates exactly what the reference validates (S/valcheck.cpp) and returns the same HRESULTs
// (S/c/ciphertext.cpp:472-577): E_INVALIDARG for a too-small buffer or an unsupported compression mode,
// COR_E_INVALIDOPERATION for malformed / invalid data, COR_E_IO for a truncated buffer.
//
// Compression: zlib comes from the system libz; Zstandard is compiled in (1.4.5, the reference's vendored version) when its
// sources are present at build time, else bound at run time from the system libzstd.so.1 (no
// headers in this image).  A Zstandard frame written here is a valid frame for the reference's decoder and vice
// versa, but the compressed BYTES are those of the system library version, not of the zstd 1.4.5 the reference
// vendors; only compr_mode none is byte-identical (tests/sealc_checks.py::wire_format).
#include "../../include/b200_sealc.h"
#include "sealc_types.h"
#include <dlfcn.h>
#include <zlib.h>

namespace
{
using namespace b200c;

const long COR_E_IO_ = (long)0x80131620L;
struct IoErr : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

enum : uint8_t
{
    COMPR_NONE = 0,
    COMPR_ZLIB = 1,
    COMPR_ZSTD = 2
};
const uint16_t kMagic = 0xA15E;
const uint8_t kVersionMajor = 4, kVersionMinor = 0; // SEAL 4.0.0 (the vendored fork)

struct Out
{
    std::vector<uint8_t> b;
    void put(const void *p, size_t n)
    {
        const uint8_t *q = (const uint8_t *)p;
        b.insert(b.end(), q, q + n);
    }
    void u64v(u64 v) { put(&v, 8); }
};
struct In
{
    const uint8_t *p;
    size_t n, pos = 0;
    In(const uint8_t *ptr, size_t size) : p(ptr), n(size) {}
    void get(void *dst, size_t len)
    {
        if (len > n - pos)
            throw IoErr("I/O error: input buffer ended unexpectedly");
        std::memcpy(dst, p + pos, len);
        pos += len;
    }
    u64 u64v()
    {
        u64 v;
        get(&v, 8);
        return v;
    }
};

// ---- Zstandard (stable API subset; prototypes restated from the public zstd.h) ----
// B200_VENDORED_ZSTD (csrc/Makefile, set when the Zstandard 1.4.5 sources the reference vendors are available at build
// time): the library is compiled INTO this shared object with hidden visibility, so compressed streams are byte-identical
// with the reference's (seal_fhe's `deterministic` test hashes them, seal_fhe/src/encryptor_decryptor.rs:919-932).
// Otherwise the system libzstd.so.1 is bound at run time: interchangeable streams, but that version's bytes.
#ifdef B200_VENDORED_ZSTD
extern "C" {
size_t ZSTD_compressBound(size_t);
size_t ZSTD_compress(void *, size_t, const void *, size_t, int);
unsigned ZSTD_isError(size_t);
void *ZSTD_createDStream(void);
size_t ZSTD_freeDStream(void *);
size_t ZSTD_initDStream(void *);
struct ZSTD_outBuffer_s;
struct ZSTD_inBuffer_s;
size_t ZSTD_decompressStream(void *, struct ZSTD_outBuffer_s *, struct ZSTD_inBuffer_s *);
unsigned ZSTD_versionNumber(void);
}
#endif
struct ZBuf
{
    const void *src;
    size_t size, pos;
};
struct ZOut
{
    void *dst;
    size_t size, pos;
};
struct Zstd
{
    size_t (*compressBound)(size_t) = nullptr;
    size_t (*compress)(void *, size_t, const void *, size_t, int) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    void *(*createDStream)() = nullptr;
    size_t (*freeDStream)(void *) = nullptr;
    size_t (*initDStream)(void *) = nullptr;
    size_t (*decompressStream)(void *, ZOut *, ZBuf *) = nullptr;
    bool ok = false;
    Zstd()
    {
#ifdef B200_VENDORED_ZSTD
        compressBound = ZSTD_compressBound;
        compress = ZSTD_compress;
        isError = ZSTD_isError;
        createDStream = ZSTD_createDStream;
        freeDStream = ZSTD_freeDStream;
        initDStream = ZSTD_initDStream;
        decompressStream = reinterpret_cast<size_t (*)(void *, ZOut *, ZBuf *)>(ZSTD_decompressStream);
        ok = true;
        return;
#endif
        void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!h)
            return;
#define B200_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(h, name))
        B200_SYM(compressBound, "ZSTD_compressBound");
        B200_SYM(compress, "ZSTD_compress");
        B200_SYM(isError, "ZSTD_isError");
        B200_SYM(createDStream, "ZSTD_createDStream");
        B200_SYM(freeDStream, "ZSTD_freeDStream");
        B200_SYM(initDStream, "ZSTD_initDStream");
        B200_SYM(decompressStream, "ZSTD_decompressStream");
#undef B200_SYM
        ok = compressBound && compress && isError && createDStream && freeDStream && initDStream && decompressStream;
    }
    static Zstd &get()
    {
        static Zstd z;
        return z;
    }
};

} // namespace
namespace b200c {
int vendored_zstd()
{
#ifdef B200_VENDORED_ZSTD
    return 1;
#else
    return 0;
#endif
}
} // namespace b200c
namespace {
bool compr_supported(uint8_t mode)
{
    return mode == COMPR_NONE || mode == COMPR_ZLIB || (mode == COMPR_ZSTD && Zstd::get().ok);
}

// Serialization::ComprSizeEstimate (S/serialization.cpp:85-110, S/util/ztools.h:62-74)
size_t compr_size_estimate(size_t in, uint8_t mode)
{
    switch (mode)
    {
    case COMPR_NONE: return in;
    case COMPR_ZLIB: return in + (in >> 12) + (in >> 14) + (in >> 25) + 17;
    case COMPR_ZSTD:
        if (Zstd::get().ok)
            return in + (in >> 8) + (in < ((size_t)128 << 10) ? ((((size_t)128 << 10) - in) >> 11) : 0);
        /* fall through */
    default: throw InvalidArg("unsupported compression mode");
    }
}

std::vector<uint8_t> deflate_bytes(const std::vector<uint8_t> &raw, uint8_t mode)
{
    if (mode == COMPR_ZLIB)
    {
        uLongf cap = compressBound((uLong)raw.size());
        std::vector<uint8_t> out(cap);
        if (compress2(out.data(), &cap, raw.data(), (uLong)raw.size(), Z_DEFAULT_COMPRESSION) != Z_OK)
            throw LogicErr("zlib compression failed");
        out.resize(cap);
        return out;
    }
    Zstd &z = Zstd::get();
    std::vector<uint8_t> out(z.compressBound(raw.size()));
    const size_t r = z.compress(out.data(), out.size(), raw.data(), raw.size(), 3 /* ZSTD_CLEVEL_DEFAULT */);
    if (z.isError(r))
        throw LogicErr("Zstandard compression failed");
    out.resize(r);
    return out;
}

std::vector<uint8_t> inflate_bytes(const uint8_t *src, size_t len, uint8_t mode)
{
    std::vector<uint8_t> out;
    std::vector<uint8_t> chunk((size_t)1 << 18);
    if (mode == COMPR_ZLIB)
    {
        z_stream zs;
        std::memset(&zs, 0, sizeof(zs));
        if (inflateInit(&zs) != Z_OK)
            throw LogicErr("stream decompression failed");
        zs.next_in = const_cast<Bytef *>(src);
        zs.avail_in = (uInt)len;
        int rc = Z_OK;
        while (rc != Z_STREAM_END)
        {
            zs.next_out = chunk.data();
            zs.avail_out = (uInt)chunk.size();
            rc = inflate(&zs, Z_NO_FLUSH);
            if (rc != Z_OK && rc != Z_STREAM_END)
            {
                inflateEnd(&zs);
                throw LogicErr("stream decompression failed");
            }
            out.insert(out.end(), chunk.data(), chunk.data() + (chunk.size() - zs.avail_out));
            if (rc == Z_OK && zs.avail_in == 0 && zs.avail_out != 0)
            { // ran out of input before the end of the stream
                inflateEnd(&zs);
                throw LogicErr("stream decompression failed");
            }
        }
        inflateEnd(&zs);
        return out;
    }
    Zstd &z = Zstd::get();
    void *ds = z.createDStream();
    if (!ds)
        throw LogicErr("stream decompression failed");
    z.initDStream(ds);
    ZBuf in{ src, len, 0 };
    while (in.pos < in.size)
    {
        ZOut o{ chunk.data(), chunk.size(), 0 };
        const size_t r = z.decompressStream(ds, &o, &in);
        if (z.isError(r))
        {
            z.freeDStream(ds);
            throw LogicErr("stream decompression failed");
        }
        out.insert(out.end(), chunk.data(), chunk.data() + o.pos);
    }
    z.freeDStream(ds);
    return out;
}

#pragma pack(push, 1)
struct Header
{
    uint16_t magic = kMagic;
    uint8_t header_size = 0x10;
    uint8_t major = kVersionMajor, minor = kVersionMinor;
    uint8_t compr = COMPR_NONE;
    uint16_t reserved = 0;
    uint64_t size = 0;
};
#pragma pack(pop)
static_assert(sizeof(Header) == 16, "SEALHeader is 16 bytes");

// Serialization::Save (S/serialization.cpp:236-339): `members` writes the raw members
template <class F>
void save_object(Out &out, uint8_t mode, F members)
{
    if (!compr_supported(mode))
        throw InvalidArg("unsupported compression mode");
    Out raw;
    members(raw);
    Header h;
    h.compr = mode;
    if (mode == COMPR_NONE)
    {
        h.size = sizeof(Header) + raw.b.size();
        out.put(&h, sizeof(h));
        out.put(raw.b.data(), raw.b.size());
        return;
    }
    std::vector<uint8_t> z = deflate_bytes(raw.b, mode);
    h.size = sizeof(Header) + z.size();
    out.put(&h, sizeof(h));
    out.put(z.data(), z.size());
}

struct Version
{
    int major, minor;
};

// Serialization::Load (S/serialization.cpp:341-446) incl. IsCompatibleVersion / IsValidHeader (serialization.h:140-190)
template <class F>
size_t load_object(In &in, F members)
{
    const size_t start = in.pos;
    Header h;
    in.get(&h, sizeof(h));
    const bool compat = (h.major == kVersionMajor && h.minor == kVersionMinor) || (h.major == 3 && h.minor >= 4);
    if (!compat)
        throw LogicErr("incompatible version");
    if (h.magic != kMagic || h.header_size != 0x10 || !compr_supported(h.compr))
        throw LogicErr("loaded SEALHeader is invalid");
    const Version v{ h.major, h.minor };
    if (h.compr == COMPR_NONE)
    {
        members(in, v);
        if (h.size != in.pos - start)
            throw LogicErr("invalid data size");
        return (size_t)h.size;
    }
    if (h.size < sizeof(Header))
        throw LogicErr("invalid data size");
    const size_t csize = (size_t)(h.size - sizeof(Header));
    if (csize > in.n - in.pos)
        throw IoErr("I/O error: input buffer ended unexpectedly");
    std::vector<uint8_t> raw = inflate_bytes(in.p + in.pos, csize, h.compr);
    in.pos += csize;
    In sub(raw.data(), raw.size());
    members(sub, v);
    return (size_t)h.size;
}

// ---- DynArray<u64> ----
const size_t kDynArrayOverhead = sizeof(Header) + 8;
void save_words(Out &out, const u64 *w, size_t count)
{
    save_object(out, COMPR_NONE, [&](Out &o) {
        o.u64v(count);
        if (count)
            o.put(w, count * 8);
    });
}
std::vector<u64> load_words(In &in, size_t bound)
{
    std::vector<u64> w;
    load_object(in, [&](In &i, Version) {
        const u64 count = i.u64v();
        if (bound && count > bound)
            throw LogicErr("unexpected size");
        if (count > (i.n - i.pos) / 8)
            throw IoErr("I/O error: input buffer ended unexpectedly");
        w.resize((size_t)count);
        if (count)
            i.get(w.data(), (size_t)count * 8);
    });
    return w;
}

// ---- UniformRandomGeneratorInfo (S/randomgen.cpp:99-160) ----
const size_t kPrngInfoBytes = sizeof(Header) + 1 + 64;
struct PrngInfo
{
    uint8_t type = 0; // 1 = blake2xb, 2 = shake256
    b200::PrngSeed seed{};
};
void save_prng_info(Out &out, const PrngInfo &pi)
{
    save_object(out, COMPR_NONE, [&](Out &o) {
        o.put(&pi.type, 1);
        o.put(pi.seed.data(), 64);
    });
}
PrngInfo load_prng_info(In &in)
{
    PrngInfo pi;
    load_object(in, [&](In &i, Version) {
        i.get(&pi.type, 1);
        if (pi.type > 2)
            throw LogicErr("prng_type is invalid");
        i.get(pi.seed.data(), 64);
    });
    return pi;
}

// ---- Ciphertext ----
bool has_seed_marker(const Ciphertext_ &ct)
{
    return ct.words() && ct.size == 2 && ct.host[(size_t)(ct.n * ct.k)] == ~(u64)0;
}
size_t ct_members_size(const Ciphertext_ &ct)
{
    ct.sync_host();
    const size_t data = has_seed_marker(ct) ? kDynArrayOverhead + ct.words() / 2 * 8 + kPrngInfoBytes : kDynArrayOverhead + ct.words() * 8;
    return 32 + 1 + 8 + 8 + 8 + 8 + 8 + data;
}
void ct_save_members(const Ciphertext_ &ct, Out &o)
{
    ct.sync_host();
    o.put(ct.parms_id.data(), 32);
    const uint8_t ntt = ct.is_ntt_form ? 1 : 0;
    o.put(&ntt, 1);
    o.u64v(ct.size);
    o.u64v(ct.n);
    o.u64v(ct.k);
    o.put(&ct.scale, 8);
    o.u64v(ct.correction_factor);
    if (has_seed_marker(ct))
    { // the PRNG info lives right after the marker word of the second polynomial (S/ciphertext.cpp:204-224)
        const size_t half = ct.words() / 2;
        In info((const uint8_t *)(ct.host.data() + half + 1), (half - 1) * 8);
        PrngInfo pi = load_prng_info(info);
        save_words(o, ct.host.data(), half);
        save_prng_info(o, pi);
    }
    else
        save_words(o, ct.host.data(), ct.words());
}

// is_metadata_valid_for(Ciphertext) (S/valcheck.cpp:67-126)
bool ct_metadata_valid(Context_ *c, const Ciphertext_ &ct, bool allow_pure_key_levels)
{
    if (!c->parameters_set)
        return false;
    const int lv = c->level_of(ct.parms_id);
    if (lv < 0)
        return false;
    if (!allow_pure_key_levels && lv < c->first_level)
        return false;
    if (ct.k != (u64)c->level_k[lv] || ct.n != c->parms.n)
        return false;
    if ((ct.size < 2 && ct.size != 0) || ct.size > 16) // SEAL_CIPHERTEXT_SIZE_MIN / _MAX (S/util/defines.h)
        return false;
    return ct.scale == 1.0 && ct.correction_factor == 1;
}
bool residues_in_range(Context_ *c, const u64 *w, size_t polys, size_t k)
{
    const size_t n = c->parms.n;
    for (size_t p = 0; p < polys; p++)
        for (size_t r = 0; r < k; r++)
        {
            const u64 q = c->parms.coeff[r];
            const u64 *x = w + (p * k + r) * n;
            for (size_t i = 0; i < n; i++)
                if (x[i] >= q)
                    return false;
        }
    return true;
}

// Ciphertext::load_members (S/ciphertext.cpp:240-372); fills `dst` host-side
void ct_load_members(Context_ *c, In &in, Version v, Ciphertext_ &dst)
{
    if (!c->parameters_set)
        throw InvalidArg("encryption parameters are not set correctly");
    Ciphertext_ t;
    in.get(t.parms_id.data(), 32);
    uint8_t ntt = 0;
    in.get(&ntt, 1);
    t.size = in.u64v();
    t.n = in.u64v();
    t.k = in.u64v();
    in.get(&t.scale, 8);
    t.correction_factor = 1;
    if (v.major == 4)
        t.correction_factor = in.u64v();
    t.is_ntt_form = ntt != 0;
    if (!ct_metadata_valid(c, t, true))
        throw LogicErr("ciphertext data is invalid");
    const size_t total = t.words();
    t.host = load_words(in, total);
    const size_t seeded = (size_t)(t.n * t.k);
    if (t.host.size() == seeded)
    { // seeded ciphertext: expand the second polynomial from the stored PRNG seed (S/ciphertext.cpp:118-151,325-352)
        PrngInfo pi;
        if (v.major == 4 || (v.major == 3 && v.minor >= 6))
            pi = load_prng_info(in);
        else if (v.major == 3 && v.minor >= 4)
        {
            pi.type = 1;
            in.get(pi.seed.data(), 64);
        }
        else
            throw LogicErr("incompatible version");
        if (pi.type != 1)
            throw LogicErr("unsupported prng_type");
        if (v.major == 3 && v.minor < 6)
            throw LogicErr("unsupported legacy seeded ciphertext"); // 3.4/3.5 used different uniform samplers
        t.host.resize(total);
        b200::Blake2xbPrng prng(pi.seed);
        std::vector<u64> mods(c->parms.coeff.begin(), c->parms.coeff.begin() + t.k);
        b200::sample_poly_uniform(prng, (size_t)t.n, mods, t.host.data() + seeded);
    }
    if (t.host.size() != total)
        throw LogicErr("ciphertext data is invalid");
    t.host_valid = true;
    dst.release_dev();
    dst.parms_id = t.parms_id;
    dst.is_ntt_form = t.is_ntt_form;
    dst.size = t.size;
    dst.n = t.n;
    dst.k = t.k;
    dst.scale = t.scale;
    dst.correction_factor = t.correction_factor;
    dst.host = std::move(t.host);
    dst.host_valid = true;
    dst.dev_valid = false;
}
// is_valid_for(Ciphertext) = buffer + metadata (data levels only) + residue ranges (S/valcheck.cpp:207-216,296-328)
void ct_check_valid(Context_ *c, const Ciphertext_ &ct)
{
    if (!ct_metadata_valid(c, ct, false) || !residues_in_range(c, ct.host.data(), (size_t)ct.size, (size_t)ct.k))
        throw LogicErr("ciphertext data is invalid");
}
// is_valid_for(PublicKey) (S/valcheck.cpp:136-146,361-394)
void pk_check_valid(Context_ *c, const Ciphertext_ &ct)
{
    if (!ct_metadata_valid(c, ct, true) || !ct.is_ntt_form || ct.parms_id != c->ids[0] || ct.size != 2 ||
        !residues_in_range(c, ct.host.data(), 2, (size_t)ct.k))
        throw LogicErr("PublicKey data is invalid");
}

// ---- Plaintext ----
size_t pt_members_size(const Plaintext_ &p) { return 32 + 8 + 8 + kDynArrayOverhead + p.coeffs.size() * 8; }
void pt_save_members(const Plaintext_ &p, Out &o)
{
    o.put(p.parms_id.data(), 32);
    o.u64v(p.coeffs.size());
    o.put(&p.scale, 8);
    save_words(o, p.coeffs.data(), p.coeffs.size());
}
// is_metadata_valid_for(Plaintext) (S/valcheck.cpp:20-65)
bool pt_metadata_valid(Context_ *c, const Plaintext_ &p, size_t coeff_count, bool allow_pure_key_levels)
{
    if (!c->parameters_set)
        return false;
    if (p.parms_id != kZeroId)
    {
        const int lv = c->level_of(p.parms_id);
        if (lv < 0 || (!allow_pure_key_levels && lv < c->first_level))
            return false;
        return (size_t)c->level_k[lv] * c->parms.n == coeff_count;
    }
    return coeff_count <= c->parms.n;
}
void pt_load_members(Context_ *c, In &in, Version, Plaintext_ &dst)
{
    if (!c->parameters_set)
        throw InvalidArg("encryption parameters are not set correctly");
    Plaintext_ t;
    in.get(t.parms_id.data(), 32);
    const u64 count = in.u64v();
    in.get(&t.scale, 8);
    if (!pt_metadata_valid(c, t, (size_t)count, true))
        throw LogicErr("plaintext data is invalid");
    t.coeffs = load_words(in, (size_t)count);
    if (t.coeffs.size() != count)
        throw LogicErr("plaintext data is invalid");
    dst = std::move(t);
}
// is_valid_for(Plaintext) (S/valcheck.cpp:246-294)
void pt_check_valid(Context_ *c, const Plaintext_ &p)
{
    bool ok = pt_metadata_valid(c, p, p.coeffs.size(), false);
    if (ok && p.parms_id != kZeroId)
        ok = residues_in_range(c, p.coeffs.data(), 1, (size_t)c->level_k[c->level_of(p.parms_id)]);
    else if (ok)
        for (u64 x : p.coeffs)
            ok = ok && x < c->parms.plain;
    if (!ok)
        throw LogicErr("Plaintext data is invalid");
}
// is_valid_for(SecretKey) (S/valcheck.cpp:128-134,330-359)
void sk_check_valid(Context_ *c, const Plaintext_ &p)
{
    if (!pt_metadata_valid(c, p, p.coeffs.size(), true) || p.parms_id != c->ids[0] ||
        !residues_in_range(c, p.coeffs.data(), 1, c->parms.coeff.size()))
        throw LogicErr("SecretKey data is invalid");
}

// ---- KSwitchKeys ----
size_t ksk_members_size(const KSwitchKeys_ &k)
{
    size_t s = 32 + 8 + 8 * k.keys.size();
    for (auto &l : k.keys)
        for (auto *pk : l)
            s += sizeof(Header) + ct_members_size(pk->data);
    return s;
}
void ksk_save_members(const KSwitchKeys_ &k, Out &o)
{
    o.put(k.parms_id.data(), 32);
    o.u64v(k.keys.size());
    for (auto &l : k.keys)
    {
        o.u64v(l.size());
        for (auto *pk : l)
            save_object(o, COMPR_NONE, [&](Out &oo) { ct_save_members(pk->data, oo); });
    }
}
void ksk_load_members(Context_ *c, In &in, Version, KSwitchKeys_ &dst)
{
    if (!c->parameters_set)
        throw InvalidArg("encryption parameters are not set correctly");
    KSwitchKeys_ t;
    in.get(t.parms_id.data(), 32);
    const u64 dim1 = in.u64v();
    if (dim1 > (in.n - in.pos) / 8)
        throw IoErr("I/O error: input buffer ended unexpectedly");
    t.keys.resize((size_t)dim1);
    for (size_t i = 0; i < dim1; i++)
    {
        const u64 dim2 = in.u64v();
        for (u64 j = 0; j < dim2; j++)
        {
            std::unique_ptr<PublicKey_> pk(new PublicKey_());
            load_object(in, [&](In &ii, Version vv) { ct_load_members(c, ii, vv, pk->data); });
            t.keys[i].push_back(pk.release());
        }
    }
    dst.clear();
    dst.parms_id = t.parms_id;
    dst.keys.swap(t.keys);
}
// is_valid_for(KSwitchKeys) (S/valcheck.cpp:148-178,396-420)
void ksk_check_valid(Context_ *c, const KSwitchKeys_ &k)
{
    bool ok = c->parameters_set && k.parms_id == c->ids[0];
    const size_t decomp = ok ? (size_t)c->level_k[c->first_level] : 0;
    for (auto &l : k.keys)
    {
        if (!ok)
            break;
        if (!l.empty() && l.size() != decomp)
            ok = false;
        for (auto *pk : l)
        {
            const Ciphertext_ &ct = pk->data;
            ok = ok && ct_metadata_valid(c, ct, true) && ct.is_ntt_form && ct.parms_id == c->ids[0] && ct.size == 2 &&
                 ct.host.size() == ct.words() && residues_in_range(c, ct.host.data(), 2, (size_t)ct.k);
        }
    }
    if (!ok)
        throw LogicErr("KSwitchKeys data is invalid");
}

template <class F>
long io_guard(F f)
{
    try
    {
        f();
        return S_OK_;
    }
    catch (const InvalidArg &)
    {
        return E_INVALIDARG_;
    }
    catch (const LogicErr &)
    {
        return COR_E_INVALIDOPERATION_;
    }
    catch (const IoErr &)
    {
        return COR_E_IO_;
    }
    catch (const std::bad_alloc &)
    {
        return E_OUTOFMEMORY_;
    }
    catch (const std::runtime_error &)
    {
        return COR_E_IO_;
    }
    catch (...)
    {
        return E_UNEXPECTED_;
    }
}

template <class SizeF>
long save_size_impl(uint8_t mode, int64_t *result, SizeF members_size)
{
    return io_guard([&] { *result = (int64_t)(sizeof(Header) + compr_size_estimate(members_size(), mode)); });
}
template <class SaveF>
long save_impl(uint8_t *outptr, uint64_t size, uint8_t mode, int64_t *out_bytes, SaveF members)
{
    return io_guard([&] {
        if (size < sizeof(Header))
            throw InvalidArg("insufficient size");
        Out out;
        save_object(out, mode, members);
        if (out.b.size() > size)
            throw IoErr("I/O error: output buffer is too small");
        std::memcpy(outptr, out.b.data(), out.b.size());
        *out_bytes = (int64_t)out.b.size();
    });
}
template <class LoadF>
long load_impl(uint8_t *inptr, uint64_t size, int64_t *in_bytes, LoadF members)
{
    return io_guard([&] {
        if (size < sizeof(Header))
            throw InvalidArg("insufficient size");
        In in(inptr, (size_t)size);
        *in_bytes = (int64_t)load_object(in, members);
    });
}

} // namespace

extern "C" {

// 1 when Zstandard 1.4.5 (the version the reference vendors) is compiled into this library: compressed streams are then
// byte-identical with the reference's; 0 when the system libzstd is bound at run time (interchangeable streams only)
long B200_VendoredZstd(void) { return b200c::vendored_zstd(); }

// ---- Ciphertext (S/c/ciphertext.cpp:472-577) ----
long Ciphertext_SaveSize(void *p, uint8_t compr_mode, int64_t *result)
{
    NULLRET(p);
    NULLRET(result);
    return save_size_impl(compr_mode, result, [&] { return ct_members_size(*(Ciphertext_ *)p); });
}
long Ciphertext_Save(void *p, uint8_t *outptr, uint64_t size, uint8_t compr_mode, int64_t *out_bytes)
{
    NULLRET(p);
    NULLRET(outptr);
    NULLRET(out_bytes);
    return save_impl(outptr, size, compr_mode, out_bytes, [&](Out &o) { ct_save_members(*(Ciphertext_ *)p, o); });
}
long Ciphertext_UnsafeLoad(void *p, void *context, uint8_t *inptr, uint64_t size, int64_t *in_bytes)
{
    NULLRET(context);
    NULLRET(p);
    NULLRET(inptr);
    NULLRET(in_bytes);
    return load_impl(inptr, size, in_bytes, [&](In &i, Version v) { ct_load_members((Context_ *)context, i, v, *(Ciphertext_ *)p); });
}
long Ciphertext_Load(void *p, void *context, uint8_t *inptr, uint64_t size, int64_t *in_bytes)
{
    NULLRET(context);
    NULLRET(p);
    NULLRET(inptr);
    NULLRET(in_bytes);
    Ciphertext_ tmp;
    long hr = load_impl(inptr, size, in_bytes, [&](In &i, Version v) {
        ct_load_members((Context_ *)context, i, v, tmp);
        ct_check_valid((Context_ *)context, tmp);
    });
    if (!hr)
        ((Ciphertext_ *)p)->assign(tmp);
    return hr;
}

// ---- Plaintext (S/c/plaintext.cpp) ----
long Plaintext_SaveSize(void *p, uint8_t compr_mode, int64_t *result)
{
    NULLRET(p);
    NULLRET(result);
    return save_size_impl(compr_mode, result, [&] { return pt_members_size(*(Plaintext_ *)p); });
}
long Plaintext_Save(void *p, uint8_t *outptr, uint64_t size, uint8_t compr_mode, int64_t *out_bytes)
{
    NULLRET(p);
    NULLRET(outptr);
    NULLRET(out_bytes);
    return save_impl(outptr, size, compr_mode, out_bytes, [&](Out &o) { pt_save_members(*(Plaintext_ *)p, o); });
}
long Plaintext_UnsafeLoad(void *p, void *context, uint8_t *inptr, uint64_t size, int64_t *in_bytes)
{
    NULLRET(context);
    NULLRET(p);
    NULLRET(inptr);
    NULLRET(in_bytes);
    return load_impl(inptr, size, in_bytes, [&](In &i, Version v) { pt_load_members((Context_ *)context, i, v, *(Plaintext_ *)p); });
}
long Plaintext_Load(void *p, void *context, uint8_t *inptr, uint64_t size, int64_t *in_bytes)
{
    NULLRET(context);
    NULLRET(p);
    NULLRET(inptr);
    NULLRET(in_bytes);
    Plaintext_ tmp;
    long hr = load_impl(inptr, size, in_bytes, [&](In &i, Version v) {
        pt_load_members((Context_ *)context, i, v, tmp);
        pt_check_valid((Context_ *)context, tmp);
    });
    if (!hr)
        *(Plaintext_ *)p = std::move(tmp);
    return hr;
}

// ---- PublicKey (S/c/publickey.cpp) ----
long PublicKey_SaveSize(void *p, uint8_t compr_mode, int64_t *result)
{
    NULLRET(p);
    NULLRET(result);
    return save_size_impl(compr_mode, result, [&] { return ct_members_size(((PublicKey_ *)p)->data); });
}
long PublicKey_Save(void *p, uint8_t *outptr, uint64_t size, uint8_t compr_mode, int64_t *out_bytes)
{
    NULLRET(p);
    NULLRET(outptr);
    NULLRET(out_bytes);
    return save_impl(outptr, size, compr_mode, out_bytes, [&](Out &o) { ct_save_members(((PublicKey_ *)p)->data, o); });
}
long PublicKey_UnsafeLoad(void *p, void *context, uint8_t *inptr, uint64_t size, int64_t *in_bytes)
{
    NULLRET(context);
    NULLRET(p);
    NULLRET(inptr);
    NULLRET(in_bytes);
    return load_impl(inptr, size, in_bytes,
                     [&](In &i, Version v) { ct_load_members((Context_ *)context, i, v, ((PublicKey_ *)p)->data); });
}
long PublicKey_Load(void *p, void *context, uint8_t *inptr, uint64_t size, int64_t *in_bytes)
{
    NULLRET(context);
    NULLRET(p);
    NULLRET(inptr);
    NULLRET(in_bytes);
    Ciphertext_ tmp;
    long hr = load_impl(inptr, size, in_bytes, [&](In &i, Version v) {
        ct_load_members((Context_ *)context, i, v, tmp);
        pk_check_valid((Context_ *)context, tmp);
    });
    if (!hr)
        ((PublicKey_ *)p)->data.assign(tmp);
    return hr;
}

// ---- SecretKey (S/c/secretkey.cpp) ----
long SecretKey_SaveSize(void *p, uint8_t compr_mode, int64_t *result)
{
    NULLRET(p);
    NULLRET(result);
    return save_size_impl(compr_mode, result, [&] { return pt_members_size(((SecretKey_ *)p)->data); });
}
long SecretKey_Save(void *p, uint8_t *outptr, uint64_t size, uint8_t compr_mode, int64_t *out_bytes)
{
    NULLRET(p);
    NULLRET(outptr);
    NULLRET(out_bytes);
    return save_impl(outptr, size, compr_mode, out_bytes, [&](Out &o) { pt_save_members(((SecretKey_ *)p)->data, o); });
}
long SecretKey_UnsafeLoad(void *p, void *context, uint8_t *inptr, uint64_t size, int64_t *in_bytes)
{
    NULLRET(context);
    NULLRET(p);
    NULLRET(inptr);
    NULLRET(in_bytes);
    return load_impl(inptr, size, in_bytes,
                     [&](In &i, Version v) { pt_load_members((Context_ *)context, i, v, ((SecretKey_ *)p)->data); });
}
long SecretKey_Load(void *p, void *context, uint8_t *inptr, uint64_t size, int64_t *in_bytes)
{
    NULLRET(context);
    NULLRET(p);
    NULLRET(inptr);
    NULLRET(in_bytes);
    Plaintext_ tmp;
    long hr = load_impl(inptr, size, in_bytes, [&](In &i, Version v) {
        pt_load_members((Context_ *)context, i, v, tmp);
        sk_check_valid((Context_ *)context, tmp);
    });
    if (!hr)
        ((SecretKey_ *)p)->data = std::move(tmp);
    return hr;
}

// ---- KSwitchKeys / RelinKeys / GaloisKeys (S/c/kswitchkeys.cpp) ----
long KSwitchKeys_SaveSize(void *p, uint8_t compr_mode, int64_t *result)
{
    NULLRET(p);
    NULLRET(result);
    return save_size_impl(compr_mode, result, [&] { return ksk_members_size(*(KSwitchKeys_ *)p); });
}
long KSwitchKeys_Save(void *p, uint8_t *outptr, uint64_t size, uint8_t compr_mode, int64_t *out_bytes)
{
    NULLRET(p);
    NULLRET(outptr);
    NULLRET(out_bytes);
    return save_impl(outptr, size, compr_mode, out_bytes, [&](Out &o) { ksk_save_members(*(KSwitchKeys_ *)p, o); });
}
long KSwitchKeys_UnsafeLoad(void *p, void *context, uint8_t *inptr, uint64_t size, int64_t *in_bytes)
{
    NULLRET(context);
    NULLRET(p);
    NULLRET(inptr);
    NULLRET(in_bytes);
    return load_impl(inptr, size, in_bytes, [&](In &i, Version v) { ksk_load_members((Context_ *)context, i, v, *(KSwitchKeys_ *)p); });
}
long KSwitchKeys_Load(void *p, void *context, uint8_t *inptr, uint64_t size, int64_t *in_bytes)
{
    NULLRET(context);
    NULLRET(p);
    NULLRET(inptr);
    NULLRET(in_bytes);
    KSwitchKeys_ tmp;
    long hr = load_impl(inptr, size, in_bytes, [&](In &i, Version v) {
        ksk_load_members((Context_ *)context, i, v, tmp);
        ksk_check_valid((Context_ *)context, tmp);
    });
    if (!hr)
    {
        auto *dst = (KSwitchKeys_ *)p;
        dst->clear();
        dst->parms_id = tmp.parms_id;
        dst->keys.swap(tmp.keys);
    }
    return hr;
}

} // extern "C"
