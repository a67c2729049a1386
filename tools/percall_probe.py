"""Developer tool: throughput of the per-handle FFI path (one Evaluator_Multiply + Evaluator_Relinearize per ciphertext pair,
as seal_fhe issues them), single-threaded and from several threads, next to the batched seam."""
import os, sys, time, threading, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunscreen_b200 import seal_fhe as s
params = (s.BfvEncryptionParametersBuilder().set_poly_modulus_degree(8192)
          .set_coefficient_modulus(s.CoefficientModulus.bfv_default(8192, s.SecurityLevel.TC128))
          .set_plain_modulus(s.PlainModulus.batching(8192, 20)).build())
ctx = s.Context(params, True, s.SecurityLevel.TC128)
gen = s.KeyGenerator(ctx)
enc = s.Encryptor.with_public_and_secret_key(ctx, gen.create_public_key(), gen.secret_key())
encoder = s.BFVEncoder(ctx)
ev = s.BFVEvaluator(ctx)
rk = gen.create_relinearization_keys()
cts = [enc.encrypt(encoder.encode_unsigned([i + 1] * 8)) for i in range(8)]
def work(n, i0=0):
    out = None
    for i in range(n):
        out = ev.relinearize(ev.multiply(cts[(i0 + i) % 8], cts[(i0 + i + 3) % 8]), rk)
    return out
work(16)
N = 256
t0 = time.perf_counter(); r = work(N); r.get_data(0); dt = time.perf_counter() - t0
print(f"1 thread : {N/dt:8.0f} mul+relin/s ({dt/N*1e6:.0f} us per pair)")
for T in (2, 4, 8):
    ths = [threading.Thread(target=work, args=(N // T, k)) for k in range(T)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    print(f"{T} threads: {N/dt:8.0f} mul+relin/s")
