// tcgen05 (5th-gen tensor core) kernels: bf16 operands, fp32 accumulation in TMEM.
//
// same_block_tc_kernel: one whole PreActFixupResBlock in mode 'same' at C = 64, 32x32
// (reference: vq_ae/layers/conv_block.py:196-216; the trunk pre_enc_layers / post_enc_layers,
// vq_ae/model.py:150-153,240-263, and the post-'down' blocks of the last pyramid level) as three
// chained implicit GEMMs per half-image tile, all intermediates on chip:
//
//   A1 = bf16(elu(x + b1a) + b1b)          18 rows x 32 px (16 rows + circular y-halo), from HBM
//   D1 = A1 . W1^T                          tcgen05.mma  M=128 N=64 K=64           (1x1)
//   U  = bf16(elu(D1 + b2a) + b2b)          written "padded-linear": 18 rows x 34 px, the two
//                                           extra columns hold the circular x-halo
//   D2 = sum_taps U[q + dy*34 + dx] . W2^T  9 taps x K=64; a tap is a constant 16-byte-granular
//                                           shift of the operand start address           (3x3)
//   V  = bf16(elu(D2 + b3a) + b3b)
//   D3 = V . W3^T                                                                        (1x1)
//   out = x + scale * D3 + b4               fp32, to HBM
//
// Operands use the un-swizzled K-major canonical layout [k-chunk][pixel][8 x bf16], so any
// pixel shift keeps the descriptor regular (tc_common.cuh).  W1/W3 stay resident in shared
// memory, the nine 8 KB W2 taps stream through a 4-slot ring filled by bulk async copies
// (UBLKCP) from L2.  Warps 0-7: prologue/epilogue math; warp 8: MMA issue; warp 9: W2 producer.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

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

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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

    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t ad = make_desc(smem_u32(sa) + row_shift * 16 + ks * 2 * a_lbo, a_lbo, 128);
            const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * b_lbo, b_lbo, 128);
            umma_bf16(tmem_base, ad, bd, idesc, ks > 0);
        }
        umma_commit(smem_u32(&bar));
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
// fused 'same' block, C = 64, 32 x 32 images, half-image tiles
// -----------------------------------------------------------------------------------------------
constexpr int SB_C = 64;
constexpr int SB_HW = 32;           // image height == width
constexpr int SB_TH = 16;           // image rows per tile
constexpr int SB_PW = SB_HW + 2;    // padded row pitch of U (pixels)
constexpr int SB_KCH = SB_C / 8;    // 16-byte k-chunks per pixel
constexpr int SB_XPIX = 641;        // A1 / V region: 5 M-tiles x 128 px, odd pitch
constexpr int SB_UPIX = 711;        // U region: 35 + 5*128 + 35 px, odd pitch
constexpr uint32_t SB_XLBO = SB_XPIX * 16;
constexpr uint32_t SB_ULBO = SB_UPIX * 16;
constexpr uint32_t SB_WLBO = SB_C * 16;              // weights: [k-chunk][n][8]
constexpr uint32_t SB_WTAP = SB_KCH * SB_WLBO;       // 8192 B per 64x64 bf16 matrix
constexpr int SB_RING = 4;
constexpr int SB_WORKERS = 256;
constexpr int SB_THREADS = SB_WORKERS + 64;
constexpr uint32_t SB_OFF_X = 0;
constexpr uint32_t SB_OFF_U = SB_OFF_X + SB_KCH * SB_XLBO;
constexpr uint32_t SB_OFF_W1 = SB_OFF_U + SB_KCH * SB_ULBO;
constexpr uint32_t SB_OFF_W3 = SB_OFF_W1 + SB_WTAP;
constexpr uint32_t SB_OFF_RING = SB_OFF_W3 + SB_WTAP;
constexpr uint32_t SB_OFF_BAR = SB_OFF_RING + SB_RING * SB_WTAP;
constexpr uint32_t SB_SMEM = SB_OFF_BAR + 128;

struct SameBlockArgs {
    const float* x;               // NHWC fp32 [B,32,32,64]
    float* out;                   // NHWC fp32 [B,32,32,64]
    const __nv_bfloat16* w;       // [W1 | W2 tap 0..8 | W3], each [k-chunk][n][8]
    int n_tiles;                  // B * 2
    float b1a, b1b, b2a, b2b, b3a, b3b, b4, scale;
};

__global__ void __launch_bounds__(SB_THREADS, 1) same_block_tc_kernel(SameBlockArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sX = sbase + SB_OFF_X, sU = sbase + SB_OFF_U, sW1 = sbase + SB_OFF_W1,
                   sW3 = sbase + SB_OFF_W3, sRing = sbase + SB_OFF_RING;
    const uint32_t bar_mma = sbase + SB_OFF_BAR;             // MMA phase complete
    const uint32_t bar_full = bar_mma + 8;                   // [SB_RING] tap landed
    const uint32_t bar_empty = bar_full + 8 * SB_RING;       // [SB_RING] tap consumed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SB_OFF_BAR + 8 + 16 * SB_RING);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_taps = my_tiles * 9;

    // ---- one-time setup ----
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        for (int s = 0; s < SB_RING; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 512);
    {   // resident W1 / W3 (generic-proxy copies; made visible to the async proxy below)
        const uint4* g1 = reinterpret_cast<const uint4*>(a.w);
        const uint4* g3 = reinterpret_cast<const uint4*>(a.w) + 10 * (SB_WTAP / 16);
        for (int i = tid; i < (int)(SB_WTAP / 16); i += SB_THREADS) {
            *reinterpret_cast<uint4*>(smem + SB_OFF_W1 + i * 16) = __ldg(g1 + i);
            *reinterpret_cast<uint4*>(smem + SB_OFF_W3 + i * 16) = __ldg(g3 + i);
        }
        // zero the activation regions once so never-written slack rows hold finite values
        for (int i = tid; i < (int)(SB_OFF_W1 / 16); i += SB_THREADS)
            *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = make_idesc_bf16(128, SB_C);
    const uint8_t* w2g = reinterpret_cast<const uint8_t*>(a.w) + SB_WTAP;

    int taps_issued = 0;                  // producer state (warp 9, lane 0)
    if (warp == 9 && lane == 0) {
        for (; taps_issued < SB_RING && taps_issued < total_taps; ++taps_issued) {
            const uint32_t fb = bar_full + 8 * (taps_issued % SB_RING);
            mbar_arrive_expect_tx(fb, SB_WTAP);
            bulk_g2s(sRing + (taps_issued % SB_RING) * SB_WTAP, w2g + (taps_issued % 9) * SB_WTAP,
                     SB_WTAP, fb);
        }
    }
    int taps_used = 0;                    // consumer state (warp 8, lane 0)
    uint32_t mma_phase = 0;

    // epilogue geometry of this worker thread
    const int q4 = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp >> 2) & 1;     // channel half: k-chunks 4*half .. 4*half+3
    const int row_in_tile = q4 * 32 + lane;
    const uint32_t t_lane = (uint32_t)(q4 * 32) << 16;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int img = tile >> 1, r0 = (tile & 1) * SB_TH;
        const float* ximg = a.x + (size_t)img * SB_HW * SB_HW * SB_C;
        float* oimg = a.out + (size_t)img * SB_HW * SB_HW * SB_C;

        // ---- P: A1 = bf16(elu(x + b1a) + b1b), 18 rows with circular y-halo ----
        if (warp < 8) {
            for (int id = tid; id < (SB_TH + 2) * SB_HW * SB_KCH; id += SB_WORKERS) {
                const int p = id >> 3, kc = id & 7;
                const int lr = p >> 5, col = p & 31;
                const int row = (r0 - 1 + lr) & (SB_HW - 1);
                const float4* src = reinterpret_cast<const float4*>(
                    ximg + ((size_t)row * SB_HW + col) * SB_C + kc * 8);
                const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
                uint4 o;
                o.x = pack_bf16(elu_fast(v0.x + a.b1a) + a.b1b, elu_fast(v0.y + a.b1a) + a.b1b);
                o.y = pack_bf16(elu_fast(v0.z + a.b1a) + a.b1b, elu_fast(v0.w + a.b1a) + a.b1b);
                o.z = pack_bf16(elu_fast(v1.x + a.b1a) + a.b1b, elu_fast(v1.y + a.b1a) + a.b1b);
                o.w = pack_bf16(elu_fast(v1.z + a.b1a) + a.b1b, elu_fast(v1.w + a.b1a) + a.b1b);
                *reinterpret_cast<uint4*>(smem + SB_OFF_X + kc * SB_XLBO + p * 16) = o;
            }
        }
        fence_proxy_async_smem();
        __syncthreads();

        // ---- G1: D1[t] = A1[t] . W1^T, 5 M-tiles (4.5 needed) ----
        if (warp == 8) {
            if (lane == 0) {
                tc_fence_after_sync();
                for (int t = 0; t < 5; ++t)
#pragma unroll
                    for (int ks = 0; ks < SB_C / 16; ++ks)
                        umma_bf16(tmem_base + t * SB_C,
                                  make_desc(sX + t * 128 * 16 + ks * 2 * SB_XLBO, SB_XLBO, 128),
                                  make_desc(sW1 + ks * 2 * SB_WLBO, SB_WLBO, 128), idesc, ks > 0);
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        // ---- E1: U = bf16(elu(D1 + b2a) + b2b) -> padded-linear with x-halo columns ----
        if (warp < 8) {
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            for (int t = 0; t < 5; ++t) {
                float v[32];
                tmem_ld32(tmem_base + t_lane + t * SB_C + half * 32, v);
                tmem_ld_wait();
                const int p = t * 128 + row_in_tile;
                if (p < (SB_TH + 2) * SB_HW) {
                    const int lr = p >> 5, col = p & 31;
                    const int qu = lr * SB_PW + col + 1;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        o.x = pack_bf16(elu_fast(v[8 * j + 0] + a.b2a) + a.b2b, elu_fast(v[8 * j + 1] + a.b2a) + a.b2b);
                        o.y = pack_bf16(elu_fast(v[8 * j + 2] + a.b2a) + a.b2b, elu_fast(v[8 * j + 3] + a.b2a) + a.b2b);
                        o.z = pack_bf16(elu_fast(v[8 * j + 4] + a.b2a) + a.b2b, elu_fast(v[8 * j + 5] + a.b2a) + a.b2b);
                        o.w = pack_bf16(elu_fast(v[8 * j + 6] + a.b2a) + a.b2b, elu_fast(v[8 * j + 7] + a.b2a) + a.b2b);
                        uint8_t* dst = smem + SB_OFF_U + (half * 4 + j) * SB_ULBO + qu * 16;
                        *reinterpret_cast<uint4*>(dst) = o;
                        if (col == 0) *reinterpret_cast<uint4*>(dst + 32 * 16) = o;            // right halo
                        if (col == SB_HW - 1) *reinterpret_cast<uint4*>(dst - 32 * 16) = o;   // left halo
                    }
                }
            }
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        fence_proxy_async_smem();
        __syncthreads();

        // ---- G2: D2[t] = sum over 9 taps of U[shifted] . W2[tap]^T ----
        if (warp == 9 && lane == 0) {
            for (int i = 0; i < 9 && taps_issued < total_taps; ++i, ++taps_issued) {
                const int slot = taps_issued % SB_RING;
                mbar_wait(bar_empty + 8 * slot, ((taps_issued / SB_RING) - 1) & 1);
                mbar_arrive_expect_tx(bar_full + 8 * slot, SB_WTAP);
                bulk_g2s(sRing + slot * SB_WTAP, w2g + (taps_issued % 9) * SB_WTAP, SB_WTAP,
                         bar_full + 8 * slot);
            }
        }
        if (warp == 8) {
            if (lane == 0) {
                tc_fence_after_sync();
                for (int tap = 0; tap < 9; ++tap, ++taps_used) {
                    const int slot = taps_used % SB_RING;
                    mbar_wait(bar_full + 8 * slot, (taps_used / SB_RING) & 1);
                    tc_fence_after_sync();
                    const int shift = (tap / 3 - 1) * SB_PW + (tap % 3 - 1);
                    for (int t = 0; t < 5; ++t)
#pragma unroll
                        for (int ks = 0; ks < SB_C / 16; ++ks)
                            umma_bf16(tmem_base + t * SB_C,
                                      make_desc(sU + (SB_PW + 1 + t * 128 + shift) * 16 + ks * 2 * SB_ULBO,
                                                SB_ULBO, 128),
                                      make_desc(sRing + slot * SB_WTAP + ks * 2 * SB_WLBO, SB_WLBO, 128),
                                      idesc, (tap | ks) > 0);
                    umma_commit(bar_empty + 8 * slot);
                }
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        // ---- E2: V = bf16(elu(D2 + b3a) + b3b), interior pixels only ----
        if (warp < 8) {
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            for (int t = 0; t < 5; ++t) {
                float v[32];
                tmem_ld32(tmem_base + t_lane + t * SB_C + half * 32, v);
                tmem_ld_wait();
                const int q = SB_PW + 1 + t * 128 + row_in_tile;
                const int lr = q / SB_PW, pc = q - lr * SB_PW;
                if (lr <= SB_TH && pc >= 1 && pc <= SB_HW) {
                    const int p = (lr - 1) * SB_HW + pc - 1;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        o.x = pack_bf16(elu_fast(v[8 * j + 0] + a.b3a) + a.b3b, elu_fast(v[8 * j + 1] + a.b3a) + a.b3b);
                        o.y = pack_bf16(elu_fast(v[8 * j + 2] + a.b3a) + a.b3b, elu_fast(v[8 * j + 3] + a.b3a) + a.b3b);
                        o.z = pack_bf16(elu_fast(v[8 * j + 4] + a.b3a) + a.b3b, elu_fast(v[8 * j + 5] + a.b3a) + a.b3b);
                        o.w = pack_bf16(elu_fast(v[8 * j + 6] + a.b3a) + a.b3b, elu_fast(v[8 * j + 7] + a.b3a) + a.b3b);
                        *reinterpret_cast<uint4*>(smem + SB_OFF_X + (half * 4 + j) * SB_XLBO + p * 16) = o;
                    }
                }
            }
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        fence_proxy_async_smem();
        __syncthreads();

        // ---- G3: D3[t] = V[t] . W3^T, 4 M-tiles ----
        if (warp == 8) {
            if (lane == 0) {
                tc_fence_after_sync();
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int ks = 0; ks < SB_C / 16; ++ks)
                        umma_bf16(tmem_base + t * SB_C,
                                  make_desc(sX + t * 128 * 16 + ks * 2 * SB_XLBO, SB_XLBO, 128),
                                  make_desc(sW3 + ks * 2 * SB_WLBO, SB_WLBO, 128), idesc, ks > 0);
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        // ---- E3: out = x + scale * D3 + b4 (fp32) ----
        if (warp < 8) {
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            for (int t = 0; t < 4; ++t) {
                float v[32];
                tmem_ld32(tmem_base + t_lane + t * SB_C + half * 32, v);
                tmem_ld_wait();
                const int p = t * 128 + row_in_tile;
                const size_t off = ((size_t)(r0 + (p >> 5)) * SB_HW + (p & 31)) * SB_C + half * 32;
                const float4* xr = reinterpret_cast<const float4*>(ximg + off);
                float4* orow = reinterpret_cast<float4*>(oimg + off);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 r = __ldg(xr + j);
                    float4 o;
                    o.x = fmaf(v[4 * j + 0], a.scale, a.b4) + r.x;
                    o.y = fmaf(v[4 * j + 1], a.scale, a.b4) + r.y;
                    o.z = fmaf(v[4 * j + 2], a.scale, a.b4) + r.z;
                    o.w = fmaf(v[4 * j + 3], a.scale, a.b4) + r.w;
                    orow[j] = o;
                }
            }
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        __syncthreads();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// OIHW fp32 weights of one 'same' block -> bf16 [11 matrices][k-chunk][n][8]:
// matrix 0 = branch_conv1, 1..9 = branch_conv2 taps (ky*3+kx), 10 = branch_conv3
__global__ void __launch_bounds__(256)
pack_same_block_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                       const float* __restrict__ w3, int C, __nv_bfloat16* __restrict__ out) {
    const int per = C * C;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 11 * per) return;
    const int m = i / per, r = i % per;
    const int kc = r / (C * 8), n = (r / 8) % C, k = kc * 8 + (r % 8);
    float v;
    if (m == 0) v = w1[n * C + k];
    else if (m == 10) v = w3[n * C + k];
    else v = w2[((size_t)n * C + k) * 9 + (m - 1)];
    out[i] = __float2bfloat16_rn(v);
}

}  // namespace

int pack_same_block_bf16(const float* w1, const float* w2, const float* w3, int C, void* packed,
                         cudaStream_t stream) {
    if (!w1 || !w2 || !w3 || !packed) return VQAE_ERR_BAD_ARG;
    if (C % 8 != 0) return VQAE_ERR_UNSUPPORTED;
    const int total = 11 * C * C;
    pack_same_block_kernel<<<ceil_div_u(total, 256), 256, 0, stream>>>(
        w1, w2, w3, C, reinterpret_cast<__nv_bfloat16*>(packed));
    return check_launch();
}

// -----------------------------------------------------------------------------------------------
int tc_selftest(const void* A, int a_rows, int row_shift, const void* B, float* D,
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

int same_block_tc(const float* x, float* out, const void* w_packed, const float* scalars8,
                  int64_t B, int H, int W, int C, int sm_count, cudaStream_t stream) {
    if (!x || !out || !w_packed || !scalars8 || B <= 0) return VQAE_ERR_BAD_ARG;
    if (x == out) return VQAE_ERR_BAD_ARG;
    if (H != SB_HW || W != SB_HW || C != SB_C) return VQAE_ERR_UNSUPPORTED;
    SameBlockArgs a;
    a.x = x; a.out = out; a.w = reinterpret_cast<const __nv_bfloat16*>(w_packed);
    a.n_tiles = (int)(B * 2);
    a.b1a = scalars8[0]; a.b1b = scalars8[1]; a.b2a = scalars8[2]; a.b2b = scalars8[3];
    a.b3a = scalars8[4]; a.b3b = scalars8[5]; a.b4 = scalars8[6]; a.scale = scalars8[7];
    static bool attr_set = false;
    if (!attr_set) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(same_block_tc_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)SB_SMEM));
        attr_set = true;
    }
    const int grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;
    same_block_tc_kernel<<<grid, SB_THREADS, SB_SMEM, stream>>>(a);
    return check_launch();
}

}  // namespace vqae
