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
    X(13, 256, 2) X(13, 256, 4) X(13, 256, 6) X(13, 256, 8) X(13, 256, 14)

b200_ntt_fp_fn b200_ntt_fp_kernel(int logn, bool fwd, int nt, int var)
{
#define X(LOGN, NT, VAR)                                                                                               \
    if (logn == LOGN && nt == NT && var == VAR)                                                                        \
        return fwd ? ntt_fp_kernel<LOGN, true, NT, VAR> : ntt_fp_kernel<LOGN, false, NT, VAR>;
    B200_FP_KERNELS(X)
#undef X
    return nullptr;
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
