// ntt_fp_kernels.cu — the statically scheduled FP64-pipe NTT / INTT kernels (ntt_fp_body.cuh) as their own translation
// unit: every (size, direction, CTA width, variant) instantiation is compiled here once; b200_bfv.cu launches them through
// the function pointers returned by b200_ntt_fp_kernel().
#include "ntt_fp_kernels.h"
#include "ntt_fp_body.cuh"
#include <cuda_runtime.h>

// FP64-only statically scheduled kernel (all slots of the job use FP-capable primes)
template <int LOGN, bool FWD, int NT, int VAR>
__global__ void __launch_bounds__(NT, (LOGN <= 13 ? (NT <= 256 ? 3 : 2) : 1)) ntt_fp_kernel(const NttJob job)
{
    extern __shared__ u64 ntt_sm[];
    if constexpr ((VAR & 16) != 0)
    {
        // Streaming variant: a persistent CTA walks blockIdx.x, blockIdx.x + gridDim.x, ... ; while one polynomial's LAST pass
        // runs, the next polynomial's words are already requested (cp.async) into the shared-memory slots that pass has
        // finished with, so no CTA ever sits idle waiting for its input (tools/ntt_ablate.py: with global traffic removed the
        // same kernel needs 0.49 ms instead of 0.70 ms for 16384 transforms — that wait is what this variant hides).
        constexpr int N = 1 << LOGN;
        const long long total = job.items * (long long)job.slots;
        long long b = (long long)blockIdx.x;
        if (b >= total)
            return;
        const int tid = (int)threadIdx.x;
        double *smd = reinterpret_cast<double *>(ntt_sm);
        if (job.stagger > 0 && job.sm_count > 0)
        { // persistent CTAs would otherwise march in lockstep: all loading, then all computing
            const long long wait = (long long)(blockIdx.x / (unsigned)job.sm_count) * job.stagger;
            const long long t0 = clock64();
            while (clock64() - t0 < wait)
                ;
        }
        int slot = job.slot_major ? (int)(b / job.items) : (int)(b % job.slots);
        long long item = job.slot_major ? b - (long long)slot * job.items : b / job.slots;
        const u64 *src = ntt_src_ptr(job, item, slot);
        {
            const int ptid = ntt_pad(tid);
            constexpr int PNT = NT + (NT >> 4);
            const unsigned sbase = (unsigned)__cvta_generic_to_shared(smd + ptid);
#pragma unroll
            for (int it = 0; it < N / NT; it++)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sbase + (unsigned)(it * PNT * 8)), "l"(src + tid + it * NT) : "memory");
        }
        for (;;)
        {
            const int pidx = job.slot_prime[slot];
            const NttPrimeFp PF = job.fprimes[pidx];
            const NttPrime PI_ = job.primes[pidx];
            u64 *dst = job.dst + item * job.dst_item_stride + job.slot_dst[slot];
            const long long nb = b + (long long)gridDim.x;
            const u64 *nsrc = nullptr;
            int nslot = 0;
            long long nitem = 0;
            if (nb < total)
            {
                nslot = job.slot_major ? (int)(nb / job.items) : (int)(nb % job.slots);
                nitem = job.slot_major ? nb - (long long)nslot * job.items : nb / job.slots;
                nsrc = ntt_src_ptr(job, nitem, nslot);
            }
            NttFpStaticPass<LOGN, NT, FWD, 0, VAR>::run(job, PF, PI_, src, dst, smd, tid, item, slot, nsrc);
            if (!nsrc)
                break;
            b = nb;
            slot = nslot;
            item = nitem;
            src = nsrc;
        }
        return;
    }
    const long long block = (long long)blockIdx.x;
    // slot-major order: CTAs that run at the same time work on the same prime, so the early-pass twiddles stay in L1
    const int slot = job.slot_major ? (int)(block / job.items) : (int)(block % job.slots);
    const long long item = job.slot_major ? block - (long long)slot * job.items : block / job.slots;
    const int pidx = job.slot_prime[slot];
    const NttPrimeFp PF = job.fprimes[pidx];
    const NttPrime PI_ = job.primes[pidx];
    const u64 *src = ntt_src_ptr(job, item, slot);
    u64 *dst = job.dst + item * job.dst_item_stride + job.slot_dst[slot];
    if (job.timeline && threadIdx.x == 0)
    {
        unsigned long long t;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        job.timeline[block * 8] = smid;
        job.timeline[block * 8 + 1] = t;
    }
    if (job.prefetch_dist > 0 && !job.tensor_mode)
    {
        // pull the polynomial that the CTA one wave later will transform into L2 (its pass-1 loads then hit L2)
        const long long nb = block + job.prefetch_dist;
        if (nb < (long long)gridDim.x)
        {
            const int nslot = job.slot_major ? (int)(nb / job.items) : (int)(nb % job.slots);
            const long long nitem = job.slot_major ? nb - (long long)nslot * job.items : nb / job.slots;
            const char *np_ = reinterpret_cast<const char *>(ntt_src_ptr(job, nitem, nslot));
            constexpr int LINES = (8 << LOGN) / 128;
#pragma unroll
            for (int l = (int)threadIdx.x; l < LINES; l += NT)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(np_ + (size_t)l * 128));
        }
    }
    NttFpStaticPass<LOGN, NT, FWD, 0, VAR>::run(job, PF, PI_, src, dst, reinterpret_cast<double *>(ntt_sm), (int)threadIdx.x, item, slot);
    if (job.timeline && threadIdx.x == 0)
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        job.timeline[block * 8 + 7] = t;
    }
}


#define B200_FP_KERNELS(X)                                                                                             \
    X(12, 256, 0) X(12, 256, 1)                                                                                        \
    X(13, 256, 0) X(13, 256, 1) X(13, 512, 0) X(13, 512, 1)                                                            \
    X(14, 1024, 0) X(14, 1024, 1)                                                                                      \
    /* developer ablations (tools/ntt_ablate.py), n = 8192 throughput configuration only */                            \
    X(13, 256, 2) X(13, 256, 4) X(13, 256, 6) X(13, 256, 8) X(13, 256, 14) X(13, 256, 32) X(13, 256, 64) X(13, 256, 128) X(13, 256, 256) X(13, 256, 384) X(13, 256, 512) X(13, 256, 896) X(13, 256, 1024) X(13, 256, 1025)                                                                           \
    /* warp-private sub-transforms (+ streaming inverse) */                                                             \
    X(12, 256, 2048) X(12, 256, 2049) X(13, 512, 2048) X(13, 512, 2049) X(13, 256, 2048) X(13, 256, 2064) X(13, 256, 2049) X(14, 1024, 2048) X(14, 1024, 2049) X(14, 1024, 2064)                                             \
    /* streaming (persistent) variant */                                                                               \
    X(12, 256, 16) X(13, 256, 16) X(14, 1024, 16)

b200_ntt_fp_fn b200_ntt_fp_kernel(int logn, bool fwd, int nt, int var)
{
#define X(LOGN, NT, VAR)                                                                                               \
    if (logn == LOGN && nt == NT && var == VAR)                                                                        \
        return fwd ? ntt_fp_kernel<LOGN, true, NT, VAR> : ntt_fp_kernel<LOGN, false, NT, VAR>;
    B200_FP_KERNELS(X)
#undef X
    return nullptr;
}

int b200_ntt_fp_ctas_per_sm(b200_ntt_fp_fn fn, int nt, size_t smem)
{
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, nt, smem) != cudaSuccess || nb < 1)
        nb = 1;
    return nb;
}

int b200_ntt_fp_setup(int smem_optin)
{
    cudaError_t e;
#define X(LOGN, NT, VAR)                                                                                               \
    if ((e = cudaFuncSetAttribute(ntt_fp_kernel<LOGN, true, NT, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin)) != cudaSuccess) return (int)e; \
    if ((e = cudaFuncSetAttribute(ntt_fp_kernel<LOGN, false, NT, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin)) != cudaSuccess) return (int)e; \
    if ((e = cudaFuncSetAttribute(ntt_fp_kernel<LOGN, true, NT, VAR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)) != cudaSuccess) return (int)e; \
    if ((e = cudaFuncSetAttribute(ntt_fp_kernel<LOGN, false, NT, VAR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)) != cudaSuccess) return (int)e;
    B200_FP_KERNELS(X)
#undef X
    return 0;
}
