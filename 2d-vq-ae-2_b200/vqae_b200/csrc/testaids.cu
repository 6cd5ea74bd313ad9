// Test / measurement aids, built into libvqae_b200_testaids.so (include/vqae_b200_testaids.h) and
// kept OUT of the product library: the descriptor/TMEM self test, the tcgen05.mma issue-rate
// microbenchmarks, and the profiling hooks of the block / trunk / quantiser kernels (which call into
// libvqae_b200.so).
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"
#include "vqae_b200_testaids.h"

namespace vqae {
namespace {

using namespace tc;

// -----------------------------------------------------------------------------------------------
// self test: D[128 x 64] = A[row_shift + m][k] . B[n][k], K = 64, operands staged in the canonical
// layout with an odd pixel pitch -- validates descriptors, address shifts and the TMEM read-back
// -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tc_selftest_kernel(const __nv_bfloat16* __restrict__ A, int a_rows, int row_shift,
                   const __nv_bfloat16* __restrict__ B, float* __restrict__ D) {
    constexpr int K = 64, N = 64, KCH = K / 8;
    extern __shared__ __align__(128) uint8_t smem[];
    const int apix = a_rows | 1;                     // odd pitch, like the block kernel
    const uint32_t a_lbo = apix * 16, b_lbo = N * 16;
    uint8_t* sa = smem;
    uint8_t* sb = sa + KCH * a_lbo;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // provably warp-uniform
    const uint32_t leader = lane == 0;
    for (int i = tid; i < a_rows * KCH; i += blockDim.x) {
        const int r = i / KCH, kc = i % KCH;
        *reinterpret_cast<uint4*>(sa + kc * a_lbo + r * 16) =
            *reinterpret_cast<const uint4*>(A + (size_t)r * K + kc * 8);
    }
    for (int i = tid; i < N * KCH; i += blockDim.x) {
        const int r = i / KCH, kc = i % KCH;
        *reinterpret_cast<uint4*>(sb + kc * b_lbo + r * 16) =
            *reinterpret_cast<const uint4*>(B + (size_t)r * K + kc * 8);
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 64);
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {                   // whole warp: umma_bf16 elects the issuing lane itself
        const uint32_t idesc = make_idesc_true_bf16(128, N);
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t ad = make_desc(smem_u32(sa) + row_shift * 16 + ks * 2 * a_lbo, a_lbo, 128);
            const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * b_lbo, b_lbo, 128);
            umma_bf16(tmem_base, ad, bd, idesc, ks > 0, 1u);
        }
        umma_commit(smem_u32(&bar), 1u);
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after_sync();
    for (int h = 0; h < 2; ++h) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + h * 32, v);
        tmem_ld_wait();
        const int m = warp * 32 + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) D[(size_t)m * N + h * 32 + j] = v[j];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// -----------------------------------------------------------------------------------------------
// MMA issue-rate microbenchmark (timing only, operand contents are arbitrary): `reps` back-to-back
// tcgen05.mma of shape 128 x N x 16 (bf16) from shared memory, layout_type 0 (no swizzle, the
// canonical layout used by the block kernels) or 2 (SWIZZLE_128B).  out[0] = cycles, out[1] = reps.
// -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tc_mma_bench_kernel(int N, int layout_type, int reps, int a_stride_rows, long long* out) {
    const int nacc = layout_type >> 4 ? (layout_type >> 4) : 2;   // independent accumulators
    layout_type &= 15;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t sa = smem_u32(smem), sb = sa + 96 * 1024;
        uint64_t ad, bd;
        if (layout_type == 0) {
            ad = make_desc(sa, 641 * 16, 128);
            bd = make_desc(sb, N * 16, 128);
        } else {   // SW128 K-major: rows of 128 B, 8-row groups 1024 B apart
            ad = make_desc(sa, 16, 1024) | (2ull << 61);
            bd = make_desc(sb, 16, 1024) | (2ull << 61);
        }
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t aoff = (uint32_t)((r % 5) * a_stride_rows * (layout_type == 0 ? 16 : 128)) >> 4;
            umma_bf16(tmem_base + (r % nacc) * N, ad + aoff, bd + (uint64_t)((r & 3) * 2), idesc, 1u, 1u);
        }
        umma_commit(smem_u32(&bar), 1u);
        mbar_wait(smem_u32(&bar), 0);
        if (tid == 0) {
            out[0] = clock64() - t0;
            out[1] = reps;
        }
    }
    __syncthreads();
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// -----------------------------------------------------------------------------------------------
static int tc_selftest(const void* A, int a_rows, int row_shift, const void* B, float* D,
                cudaStream_t stream) {
    if (!A || !B || !D || a_rows < 128 || row_shift < 0 || row_shift + 128 > a_rows)
        return VQAE_ERR_BAD_ARG;
    const size_t smem = (size_t)8 * ((a_rows | 1) * 16) + 8 * 64 * 16;
    if (smem > 200 * 1024) return VQAE_ERR_UNSUPPORTED;
    VQAE_CUDA_TRY(cudaFuncSetAttribute(tc_selftest_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_selftest_kernel<<<1, 128, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(A), a_rows,
                                                 row_shift,
                                                 reinterpret_cast<const __nv_bfloat16*>(B), D);
    return check_launch();
}

static int tc_mma_bench(int N, int layout_type, int reps, int a_stride_rows, long long* out,
                 cudaStream_t stream) {
    if (!out || reps <= 0 || N < 16 || N > 256 || N % 16) return VQAE_ERR_BAD_ARG;
    const size_t smem = 160 * 1024;
    VQAE_CUDA_TRY(cudaFuncSetAttribute(tc_mma_bench_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_mma_bench_kernel<<<1, 128, smem, stream>>>(N, layout_type, reps, a_stride_rows, out);
    return check_launch();
}


}  // namespace vqae

using namespace vqae;

extern "C" {

int vqae_tc_selftest(const void* a_bf16, int a_rows, int row_shift, const void* b_bf16, float* d,
                     void* stream) {
    return tc_selftest(a_bf16, a_rows, row_shift, b_bf16, d, (cudaStream_t)stream);
}

int vqae_tc_mma_bench(int n, int layout_type, int reps, int a_stride_rows, long long* out2,
                      void* stream) {
    return tc_mma_bench(n, layout_type, reps, a_stride_rows, out2, (cudaStream_t)stream);
}

int vqae_tc_mma_bench2(int m, int n, int reps, int n_issuers, int ctas_per_sm, int mode,
                       long long* out_per_cta, void* stream) {
    return tc_mma_bench2(m, n, reps, n_issuers, ctas_per_sm, mode, out_per_cta, (cudaStream_t)stream);
}

int vqae_same_block_f16_profile(const void* x, void* out, int io_dtype, const void* w_packed,
                                const float* scalars8_host, int64_t batch, int height, int width,
                                 int c, long long* phase_clocks, void* stream) {
    int sm_count = 0;
    if (int rc = device_sm_count(&sm_count)) return rc;
    return same_block_tc(x, out, io_dtype, w_packed, scalars8_host, batch, height, width, c, sm_count,
                         phase_clocks, (cudaStream_t)stream);
}

void vqae_trunk_resident_set_profile(long long* phase_clocks) { trunk_resident_set_prof(phase_clocks); }

void vqae_quantize_tc_set_profile(long long* phase_clocks) { quantize_tc_set_prof(phase_clocks); }

}  // extern "C"
