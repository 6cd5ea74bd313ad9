// fp32-accurate PreActFixupResBlock 'same' on the tcgen05 tensor cores ("fp32tc" precision):
// every GEMM operand is SPLIT into two fp16 numbers,  v = hi + lo,  hi = f16(v),  lo = f16(v - hi)
// (22 significand bits together), and each product is evaluated as three tensor-core products
//
//      A . B  ~=  A_hi . B_hi  +  A_lo . B_hi  +  A_hi . B_lo          (lo . lo ~ 2^-22: dropped)
//
// accumulated in ONE fp32 accumulator in tensor memory -- the same device the quantiser's candidate
// filter uses (quantize_tc.cu), here applied to the three chained implicit GEMMs of the block
// (reference: vq_ae/layers/conv_block.py:196-216).  fp16 carries only 5 exponent bits, so the weight
// matrices are pre-multiplied at pack time by a power of two that moves max|w| to (2^13, 2^14]
// (pack.cu, vqae_pack_desc.premul) -- their low halves stay normal numbers down to 2^-27 of the
// largest weight -- and the accumulators are multiplied by the inverse, which is exact.  Activations
// are O(1): their low halves are normal for |v| >= 2^-3 and carry an absolute error <= 2^-25 below.
// The activation is an fp32-grade ELU (elu1_tc, common.cuh: Taylor series near zero, SFU exponential
// beyond), not the fast exponential of the reduced-precision kernels.  Purpose: the reference's index contract (vq.py:121-129: the argmin
// of fp32 distances) holds on this path outside reported near-ties, at tensor-core speed
// (tests/test_gpu_split.py pins it to the reference goldens).
//
// Structure = the tile kernel of tc_kernels.cu, simplified: tile 8 x 32 pixels + circular halo ring
// (10 x 34 = 340 padded-linear pixels, 3 M-tiles), ONE operand region reused for A1 -> U -> V (two
// planes: hi | lo), W1 / W3 resident, the W2 taps through a 3-slot bulk-copy ring at C = 64.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace vqae {
namespace {

using namespace tc;

constexpr int SP_TH = 8, SP_TW = 32, SP_PW = SP_TW + 2;
constexpr int SP_NPAD = (SP_TH + 2) * SP_PW;       // 340 padded pixels
constexpr int SP_MT = 3;                            // M-tiles of G1 / G2 (384 rows)
constexpr int SP_MT3 = SP_TH * SP_TW / 128;         // M-tiles of G3 (interior, pixel-linear)
constexpr int SP_RPIX = 455;                        // 35 + 3 * 128 + 35 = 454 pixels, odd pitch
constexpr uint32_t SP_LBO = SP_RPIX * 16;
constexpr int SP_RING = 3;                          // divides the 9 taps: a tap's slot is tap % 3, a
                                                    // compile-time constant (descriptors stay uniform)

template <int CP, int CR>
struct SplitCfg {
    static constexpr int KCH = CP / 8, KCR = CR / 8;
    static constexpr bool RING = (CP == 64);
    static constexpr int NW = CP == 64 ? 16 : (CP == 32 ? 8 : 4);
    static constexpr int WORKERS = NW * 32, THREADS = WORKERS + 64;
    static constexpr int NG = NW / 4, NC = CP / NG;                 // 16 TMEM columns per unit
    static constexpr int UCH = (CR < CP) ? KCR : NC / 8;            // real k-chunks per unit
    static constexpr uint32_t WLBO = CP * 16, WMAT = KCH * WLBO;    // one CP x CP fp16 matrix
    static constexpr uint32_t PLANE = KCH * SP_LBO;                 // hi plane; lo plane follows
    static constexpr uint32_t OFF_R = 0;
    static constexpr uint32_t OFF_W = 2 * PLANE;
    // RING: W1 hi | W1 lo | W3 hi | W3 lo | ring[s] = (tap hi | tap lo);  else: hi x 11 | lo x 11
    static constexpr uint32_t W_BYTES = (RING ? 2 * (2 + SP_RING) : 22) * WMAT;
    static constexpr uint32_t LO_OFF = RING ? WMAT : 11 * WMAT;     // lo matrix behind its hi matrix
    static constexpr uint32_t OFF_BAR = OFF_W + W_BYTES;
    static constexpr uint32_t SMEM = OFF_BAR + 128;
    static constexpr int TMEM_COLS = CP == 64 ? 256 : (CP == 32 ? 128 : 64);
    static constexpr int MIN_CTAS = CP == 64 ? 1 : (CP == 32 ? 2 : 4);
};

struct SplitArgs {
    const float* x;               // NHWC fp32 [B,H,W,CR]
    float* out;                   // NHWC fp32 [B,H,W,CR]
    const __half* w_hi;           // 11 matrices [W1 | W2 tap 0..8 | W3] (VQAE_PACK_SAME_F16, premul)
    const __half* w_lo;           // ... | VQAE_PACK_LO
    int n_tiles, H, W, tiles_x, tiles_per_img;
    float b1a, b1b, b2a, b2b, b3a, b3b, b4;
    float inv1, inv2, scale3;     // 1 / premul of W1, W2;  scale / premul of W3
};

// 8 fp32 -> hi and lo fp16 octets
__device__ __forceinline__ void split8(const float (&f)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int CP, int CR>
__global__ void __launch_bounds__(SplitCfg<CP, CR>::THREADS, SplitCfg<CP, CR>::MIN_CTAS)
same_block_split_kernel(SplitArgs a) {
    using Cfg = SplitCfg<CP, CR>;
    constexpr int KCH = Cfg::KCH, KCR = Cfg::KCR, NW = Cfg::NW, NC = Cfg::NC, UCH = Cfg::UCH;
    constexpr uint32_t WLBO = Cfg::WLBO, WMAT = Cfg::WMAT, PLANE = Cfg::PLANE, LO_OFF = Cfg::LO_OFF;
    constexpr int MMA_WARP = NW, PROD_WARP = NW + 1;

    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sR = sbase + Cfg::OFF_R, sW = sbase + Cfg::OFF_W;
    const uint32_t sW1 = sW;
    const uint32_t sW3 = Cfg::RING ? sW + 2 * WMAT : sW + 10 * WMAT;
    const uint32_t sW2 = Cfg::RING ? sW + 4 * WMAT : sW + WMAT;      // ring base / tap 0 (hi)
    const uint32_t bar_mma = sbase + Cfg::OFF_BAR;
    const uint32_t bar_full = bar_mma + 8;
    const uint32_t bar_empty = bar_full + 8 * SP_RING;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_BAR + 8 + 16 * SP_RING);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t leader = lane == 0;
    const int my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_taps = my_tiles * 9;

    if (tid == 0) {
        mbar_init(bar_mma, 1);
        for (int s = 0; s < SP_RING; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        fence_mbar_init();
    }
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
    {
        const uint4* gh = reinterpret_cast<const uint4*>(a.w_hi);
        const uint4* gl = reinterpret_cast<const uint4*>(a.w_lo);
        constexpr int MV = WMAT / 16;                                 // 16-byte pieces per matrix
        if (Cfg::RING) {
            for (int i = tid; i < MV; i += Cfg::THREADS) {
                uint4* d = reinterpret_cast<uint4*>(smem + Cfg::OFF_W) + i;
                d[0] = __ldg(gh + i);
                d[MV] = __ldg(gl + i);
                d[2 * MV] = __ldg(gh + 10 * MV + i);
                d[3 * MV] = __ldg(gl + 10 * MV + i);
            }
        } else {
            for (int i = tid; i < 11 * MV; i += Cfg::THREADS) {
                uint4* d = reinterpret_cast<uint4*>(smem + Cfg::OFF_W) + i;
                d[0] = __ldg(gh + i);
                d[11 * MV] = __ldg(gl + i);
            }
        }
        for (int i = tid; i < (int)(Cfg::OFF_W / 16); i += Cfg::THREADS)
            *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t idesc = make_idesc_bf16(128, CP);
    const uint8_t* w2h = reinterpret_cast<const uint8_t*>(a.w_hi) + WMAT;
    const uint8_t* w2l = reinterpret_cast<const uint8_t*>(a.w_lo) + WMAT;

    int taps_issued = 0;                  // producer state (PROD_WARP lane 0, RING only)
    auto issue_tap = [&](int n) {
        const int slot = n % SP_RING;
        const uint32_t fb = bar_full + 8 * slot;
        mbar_arrive_expect_tx(fb, 2 * WMAT);
        bulk_g2s(sW2 + slot * 2 * WMAT, w2h + (n % 9) * WMAT, WMAT, fb);
        bulk_g2s(sW2 + slot * 2 * WMAT + WMAT, w2l + (n % 9) * WMAT, WMAT, fb);
    };
    if (Cfg::RING && warp == PROD_WARP && lane == 0) {
        for (; taps_issued < SP_RING && taps_issued < total_taps; ++taps_issued) issue_tap(taps_issued);
    }
    int taps_used = 0;                    // consumer state (MMA_WARP, RING only)
    uint32_t mma_phase = 0;

    const int q4 = warp & 3;
    const int grp = (warp >> 2) % Cfg::NG;
    const int row_in_tile = q4 * 32 + lane;
    const uint32_t t_lane = (uint32_t)(q4 * 32) << 16;
    const uint32_t t_col = grp * NC;
    const int kc0 = grp * (NC / 8);
    constexpr int NCR = UCH * 8;
    constexpr int F4 = NCR / 4;
    constexpr int SROW = NCR + 4;
    float* stage = reinterpret_cast<float*>(smem + Cfg::OFF_R) + (warp < NW ? warp : 0) * 32 * SROW;

    const uint64_t dR = make_desc(sR, SP_LBO, 128);
    const uint64_t dW1 = make_desc(sW1, WLBO, 128);
    const uint64_t dW3 = make_desc(sW3, WLBO, 128);
    const uint64_t dW2 = make_desc(sW2, WLBO, 128);
    constexpr uint64_t A_LO = PLANE >> 4, B_LO = LO_OFF >> 4;

    // D (+)= A . B^T with split operands: lo terms first, the dominant hi . hi product last
    auto umma3 = [&](uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t accumulate) {
        umma_bf16(d_tmem, da + A_LO, db, idesc, accumulate, leader);
        umma_bf16(d_tmem, da, db + B_LO, idesc, 1u, leader);
        umma_bf16(d_tmem, da, db, idesc, 1u, leader);
    };

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int img = tile / a.tiles_per_img;
        const int trem = tile - img * a.tiles_per_img;
        const int r0 = (trem / a.tiles_x) * SP_TH, c0 = (trem % a.tiles_x) * SP_TW;
        const float* ximg = a.x + (size_t)img * a.H * a.W * CR;
        float* oimg = a.out + (size_t)img * a.H * a.W * CR;

        // ---- P: A1 = split(elu(x + b1a) + b1b) on the 10 x 34 halo'd tile (circular wrap) ----
        if (warp < NW) {
            constexpr int ITEMS = SP_NPAD * KCR;
            constexpr int PB = 4;
            for (int base = tid; base < ITEMS; base += Cfg::WORKERS * PB) {
                float vv[PB][8];
                int dst[PB];
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const int id = base + u * Cfg::WORKERS;
                    dst[u] = -1;
                    if (id < ITEMS) {
                        const int q = id / KCR, kc = id - q * KCR;
                        const int lr = q / SP_PW, lc = q - lr * SP_PW;
                        int row = r0 - 1 + lr, col = c0 - 1 + lc;
                        row = row < 0 ? row + a.H : (row >= a.H ? row - a.H : row);
                        col = col < 0 ? col + a.W : (col >= a.W ? col - a.W : col);
                        StreamIO<float>::load8(ximg + ((size_t)row * a.W + col) * CR + kc * 8, vv[u]);
                        dst[u] = kc * (int)SP_LBO + q * 16;
                    }
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    if (dst[u] >= 0) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = elu1_tc(vv[u][e] + a.b1a) + a.b1b;
                        uint4 hi, lo;
                        split8(f, hi, lo);
                        *reinterpret_cast<uint4*>(smem + Cfg::OFF_R + dst[u]) = hi;
                        *reinterpret_cast<uint4*>(smem + Cfg::OFF_R + PLANE + dst[u]) = lo;
                    }
                }
            }
        }
        fence_proxy_async_smem();
        __syncthreads();

        // ---- G1: D1 = A1 . W1^T on the 3 M-tiles of padded-linear pixels ----
        if (warp == MMA_WARP) {
            tc_fence_after_sync();
#pragma unroll
            for (int t = 0; t < SP_MT; ++t)
#pragma unroll
                for (int ks = 0; ks < CP / 16; ++ks)
                    umma3(tmem_base + t * CP, dR + (uint64_t)((t * 128 * 16 + ks * 2 * SP_LBO) >> 4),
                          dW1 + (uint64_t)((ks * 2 * WLBO) >> 4), ks > 0);
            umma_commit(bar_mma, leader);
            __syncwarp();
        }
        // ---- E1: U[q] = split(elu(D1[q] / premul1 + b2a) + b2b), written over A1 ----
        if (warp < NW) {
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            for (int t = 0; t < SP_MT; ++t) {
                float v[NC];
                tmem_ld<NC>(tmem_base + t_lane + t * CP + t_col, v);
                tmem_ld_wait();
                const int q = t * 128 + row_in_tile;
#pragma unroll
                for (int j = 0; j < UCH; ++j) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = elu1_tc(fmaf(v[8 * j + e], a.inv1, a.b2a)) + a.b2b;
                    uint4 hi, lo;
                    split8(f, hi, lo);
                    const uint32_t off = Cfg::OFF_R + (kc0 + j) * SP_LBO + q * 16;
                    *reinterpret_cast<uint4*>(smem + off) = hi;
                    *reinterpret_cast<uint4*>(smem + off + PLANE) = lo;
                }
            }
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        fence_proxy_async_smem();
        __syncthreads();

        // ---- G2: D2[q] = sum over 9 taps of U[q + dy*34 + dx] . W2[tap]^T, q from 35 ----
        if (Cfg::RING && warp == PROD_WARP && lane == 0) {
            for (int i = 0; i < 9 && taps_issued < total_taps; ++i, ++taps_issued) {
                const int slot = taps_issued % SP_RING;
                mbar_wait(bar_empty + 8 * slot, ((taps_issued / SP_RING) - 1) & 1);
                issue_tap(taps_issued);
            }
        }
        if (warp == MMA_WARP) {
            tc_fence_after_sync();
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                uint64_t dW;
                const int slot = tap % SP_RING;
                if (Cfg::RING) {
                    mbar_wait(bar_full + 8 * slot, (taps_used / SP_RING) & 1);
                    tc_fence_after_sync();
                    dW = dW2 + (uint64_t)((slot * 2 * WMAT) >> 4);
                    ++taps_used;
                } else {
                    dW = dW2 + (uint64_t)((tap * WMAT) >> 4);
                }
                const int shift = (tap / 3 - 1) * SP_PW + (tap % 3 - 1);
#pragma unroll
                for (int t = 0; t < SP_MT; ++t)
#pragma unroll
                    for (int ks = 0; ks < CP / 16; ++ks)
                        umma3(tmem_base + t * CP,
                              dR + (uint64_t)(((SP_PW + 1 + t * 128 + shift) * 16 + ks * 2 * SP_LBO) >> 4),
                              dW + (uint64_t)((ks * 2 * WLBO) >> 4), (tap | ks) > 0);
                if (Cfg::RING) umma_commit(bar_empty + 8 * slot, leader);
            }
            umma_commit(bar_mma, leader);
            __syncwarp();
        }
        // ---- E2: V[p] = split(elu(D2 / premul2 + b3a) + b3b), 8 x 32 interior, pixel-linear, over U
        if (warp < NW) {
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            for (int t = 0; t < SP_MT; ++t) {
                float v[NC];
                tmem_ld<NC>(tmem_base + t_lane + t * CP + t_col, v);
                tmem_ld_wait();
                const int q = SP_PW + 1 + t * 128 + row_in_tile;
                const int lr = q / SP_PW, pc = q - lr * SP_PW;
                if (lr <= SP_TH && pc >= 1 && pc <= SP_TW) {
                    const int p = (lr - 1) * SP_TW + pc - 1;
#pragma unroll
                    for (int j = 0; j < UCH; ++j) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = elu1_tc(fmaf(v[8 * j + e], a.inv2, a.b3a)) + a.b3b;
                        uint4 hi, lo;
                        split8(f, hi, lo);
                        const uint32_t off = Cfg::OFF_R + (kc0 + j) * SP_LBO + p * 16;
                        *reinterpret_cast<uint4*>(smem + off) = hi;
                        *reinterpret_cast<uint4*>(smem + off + PLANE) = lo;
                    }
                }
            }
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        fence_proxy_async_smem();
        __syncthreads();

        // ---- G3: D3 = V . W3^T, 2 M-tiles ----
        if (warp == MMA_WARP) {
            tc_fence_after_sync();
#pragma unroll
            for (int t = 0; t < SP_MT3; ++t)
#pragma unroll
                for (int ks = 0; ks < CP / 16; ++ks)
                    umma3(tmem_base + t * CP, dR + (uint64_t)((t * 128 * 16 + ks * 2 * SP_LBO) >> 4),
                          dW3 + (uint64_t)((ks * 2 * WLBO) >> 4), ks > 0);
            umma_commit(bar_mma, leader);
            __syncwarp();
        }
        // ---- E3: out = x + (scale / premul3) * D3 + b4, transposed through shared memory (the
        //      operand region is dead once G3 has completed) for 128-bit coalesced accesses ----
        if (warp < NW) {
            const int rsub = lane / F4, c4 = lane % F4;
            auto x_off = [&](int t, int k) {
                const int p = t * 128 + q4 * 32 + rsub + k * (32 / F4);
                return ((size_t)(r0 + (p >> 5)) * a.W + c0 + (p & 31)) * CR + kc0 * 8 + c4 * 4;
            };
            float4 xr[F4], xn[F4];
#pragma unroll
            for (int k = 0; k < F4; ++k) xr[k] = __ldg(reinterpret_cast<const float4*>(ximg + x_off(0, k)));
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
#pragma unroll
            for (int t = 0; t < SP_MT3; ++t) {
                float v[NC];
                tmem_ld<NC>(tmem_base + t_lane + t * CP + t_col, v);
                if (t + 1 < SP_MT3) {
#pragma unroll
                    for (int k = 0; k < F4; ++k)
                        xn[k] = __ldg(reinterpret_cast<const float4*>(ximg + x_off(t + 1, k)));
                }
                tmem_ld_wait();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < F4; ++j)
                    *reinterpret_cast<float4*>(stage + lane * SROW + 4 * j) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < F4; ++k) {
                    const int rr = rsub + k * (32 / F4);
                    const float4 d = *reinterpret_cast<const float4*>(stage + rr * SROW + 4 * c4);
                    float4 o;
                    o.x = fmaf(d.x, a.scale3, a.b4) + xr[k].x;
                    o.y = fmaf(d.y, a.scale3, a.b4) + xr[k].y;
                    o.z = fmaf(d.z, a.scale3, a.b4) + xr[k].z;
                    o.w = fmaf(d.w, a.scale3, a.b4) + xr[k].w;
                    *reinterpret_cast<float4*>(oimg + x_off(t, k)) = o;
                }
#pragma unroll
                for (int k = 0; k < F4; ++k) xr[k] = xn[k];
            }
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        __syncthreads();          // the staging rows are free before the next tile's A1 is written
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int CP, int CR>
int launch_split(const SplitArgs& a, int sm_count, cudaStream_t stream) {
    using Cfg = SplitCfg<CP, CR>;
    auto kern = same_block_split_kernel<CP, CR>;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)Cfg::SMEM));
        attr_set.cur() = true;
    }
    const int cap = sm_count * Cfg::MIN_CTAS;
    const int grid = a.n_tiles < cap ? a.n_tiles : cap;
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(a);
    return check_launch();
}

}  // namespace

bool same_block_split_supported(int H, int W, int C) {
    return (C == 8 || C == 16 || C == 32 || C == 64) && H >= SP_TH && W >= SP_TW && H % SP_TH == 0 &&
           W % SP_TW == 0;
}

int same_block_split(const float* x, float* out, const void* w_hi, const void* w_lo,
                     const float* scalars8, const float* premul3, int64_t B, int H, int W, int C,
                     int sm_count, cudaStream_t stream) {
    if (!x || !out || !w_hi || !w_lo || !scalars8 || !premul3 || B <= 0) return VQAE_ERR_BAD_ARG;
    if (x == out) return VQAE_ERR_BAD_ARG;
    if (!same_block_split_supported(H, W, C)) return VQAE_ERR_UNSUPPORTED;
    for (int i = 0; i < 3; ++i)
        if (!(premul3[i] > 0.f)) return VQAE_ERR_BAD_ARG;
    SplitArgs a;
    a.x = x; a.out = out;
    a.w_hi = reinterpret_cast<const __half*>(w_hi);
    a.w_lo = reinterpret_cast<const __half*>(w_lo);
    a.H = H; a.W = W; a.tiles_x = W / SP_TW; a.tiles_per_img = (H / SP_TH) * a.tiles_x;
    const int64_t nt = B * a.tiles_per_img;
    if (nt > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    a.n_tiles = (int)nt;
    a.b1a = scalars8[0]; a.b1b = scalars8[1]; a.b2a = scalars8[2]; a.b2b = scalars8[3];
    a.b3a = scalars8[4]; a.b3b = scalars8[5]; a.b4 = scalars8[6];
    a.inv1 = 1.f / premul3[0]; a.inv2 = 1.f / premul3[1]; a.scale3 = scalars8[7] / premul3[2];
    switch (C) {
        case 64: return launch_split<64, 64>(a, sm_count, stream);
        case 32: return launch_split<32, 32>(a, sm_count, stream);
        case 16: return launch_split<16, 16>(a, sm_count, stream);
        case 8: return launch_split<16, 8>(a, sm_count, stream);
    }
    return VQAE_ERR_UNSUPPORTED;
}

}  // namespace vqae
