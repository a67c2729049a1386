"""Developer tool: per-kernel CUDA-event totals (B200_TRACE) of the per-handle path: 64 x (Evaluator_Multiply + Evaluator_Relinearize)
on single ciphertexts, one thread.  Shows which launches make up the ~190 us a pair takes when nothing is batched."""
import os, sys
os.environ["B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunscreen_b200 import seal_fhe as s
from sunscreen_b200.lib import B200Lib
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
for i in range(64):
    ev.relinearize(ev.multiply(cts[i % 8], cts[(i + 3) % 8]), rk)
B200Lib.default().lib.b200_trace_dump()
