// tcgen05.mma throughput microbenchmark with several issuing warps per CTA and several CTAs per SM
// (timing only, operand contents are arbitrary): tells apart a per-issue-stream floor from a
// per-SM tensor-pipe floor.  out[cta] = cycles until every issuer's MMAs have completed.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace vqae {
namespace {
using namespace tc;

__global__ void __launch_bounds__(256)
tc_mma_bench2_kernel(int M, int N, int reps, int n_issuers, int tmem_cols, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[8];
    __shared__ uint32_t tmem_base_s;
    __shared__ long long t_end[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 48 * 1024 / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    const long long t0 = clock64();
    if (warp < n_issuers) {
        const uint32_t idesc = make_idesc_bf16(M, N);
        // mode bit 2: launched as 4-CTA clusters (host side);
        // mode bit 0: every issuer reads its OWN A and B regions (else all issuers share them);
        // mode bit 1: 8-row groups of A 160 B apart (the resident kernel's column-major tile) else 128 B
        const uint32_t a_off = (mode & 1) ? warp * 5 * 1024 : 0, b_off = (mode & 1) ? warp * 2 * 1024 : 0;
        const uint32_t sa = smem_u32(smem) + a_off, sb = smem_u32(smem) + 32 * 1024 + b_off;
        const uint64_t ad = make_desc(sa, 340 * 16, (mode & 2) ? 160 : 128), bd = make_desc(sb, N * 16, 128);
        const uint32_t d = tmem_base + warp * N;
        for (int r = 0; r < reps; ++r)
            umma_bf16(d, ad + (uint64_t)(((r % 5) * 16) >> 4), bd + (uint64_t)((r & 3) * 2), idesc, 1u,
                      lane == 0);
        umma_commit(smem_u32(&bars[warp]), lane == 0);
        mbar_wait(smem_u32(&bars[warp]), 0);
        if (lane == 0) t_end[warp] = clock64() - t0;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        long long m = 0;
        for (int i = 0; i < n_issuers; ++i) m = t_end[i] > m ? t_end[i] : m;
        out[blockIdx.x] = m;
    }
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}
}  // namespace

int tc_mma_bench2(int M, int N, int reps, int n_issuers, int ctas_per_sm, int mode, long long* out,
                  cudaStream_t stream) {
    if (!out || reps <= 0 || N < 16 || N > 256 || N % 16 || (M != 64 && M != 128) || n_issuers < 1 ||
        n_issuers > 8 || ctas_per_sm < 1 || ctas_per_sm > 4)
        return VQAE_ERR_BAD_ARG;
    int cols = 32;
    while (cols < n_issuers * N) cols *= 2;
    if (cols * ctas_per_sm > 512) return VQAE_ERR_UNSUPPORTED;
    int sm_count = 0;
    if (int rc = device_sm_count(&sm_count)) return rc;
    // shared memory sized so that exactly ctas_per_sm CTAs fit on an SM
    const size_t smem = ctas_per_sm == 1 ? 160 * 1024 : (ctas_per_sm == 2 ? 100 * 1024
                        : (ctas_per_sm == 3 ? 70 * 1024 : 52 * 1024));
    VQAE_CUDA_TRY(cudaFuncSetAttribute(tc_mma_bench2_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (mode & 4) {                                      // launched as clusters of four CTAs
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((sm_count / 4) * 4 * ctas_per_sm);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 4; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        VQAE_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc_mma_bench2_kernel, M, N, reps, n_issuers, cols, mode, out));
        return check_launch();
    }
    tc_mma_bench2_kernel<<<sm_count * ctas_per_sm, 256, smem, stream>>>(M, N, reps, n_issuers, cols, mode, out);
    return check_launch();
}
}  // namespace vqae
