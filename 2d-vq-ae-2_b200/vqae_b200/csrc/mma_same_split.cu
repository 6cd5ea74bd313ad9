// fp32-ACCURATE form of mma_same.cu (precision "fp32tc"): PreActFixupResBlock 'same' at C = 8, 16 on
// warp-level MMAs with SPLIT fp16 operands -- every operand a pair hi + lo (22 significand bits), every
// product three MMAs (lo.hi + hi.lo + hi.hi), weights pre-multiplied by powers of two at pack time
// (vqae_pack_desc.premul) and the accumulators by the inverse, an fp32-grade ELU (elu1_tc, common.cuh).
// See tc_split.cu for the numerics; the structure is that of mma_same.cu (asynchronous halo'd x tile,
// stage 1 on the ring'd tile -> U, stages 2 + 3 register-chained), one M-tile per warp step.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"
#include "mma_common.cuh"
#include "tc_common.cuh"

namespace vqae {
namespace {

using namespace mma;

constexpr int SS_TW = 32, SS_PW = SS_TW + 2;
constexpr int SS_WARPS = 8, SS_THREADS = SS_WARPS * 32;

template <int C, int TH>
struct SsCfg {
    static constexpr bool K8 = (C == 8);
    static constexpr int KS = K8 ? 1 : C / 16;
    static constexpr int NT = C / 8;
    static constexpr int NPAD = (TH + 2) * SS_PW;
    static constexpr int MT1 = (NPAD + 15) / 16;
    static constexpr int MT2 = TH * SS_TW / 16;
    static constexpr int UP = K8 ? 16 : C * 2 + 16;
    static constexpr int XP = C * 4;
    static constexpr int WP = C * 2 + 16;
    static constexpr uint32_t X_BYTES = (uint32_t)(MT1 * 16) * XP;
    static constexpr uint32_t U_PLANE = (uint32_t)(MT1 * 16) * UP;           // hi plane; lo plane behind it
    static constexpr uint32_t W_SET = K8 ? 0 : 9 * C * WP;                   // W2 hi set; lo set behind it
    static constexpr uint32_t OFF_X = 0;
    static constexpr uint32_t OFF_U = OFF_X + 2 * X_BYTES;
    static constexpr uint32_t OFF_W = OFF_U + 2 * U_PLANE;
    static constexpr uint32_t OFF_BAR = OFF_W + 2 * W_SET;
    static constexpr uint32_t SMEM = OFF_BAR + 16;
    static constexpr int MIN_CTAS = 2;
};

struct SsArgs {
    const float* x;
    float* out;
    const __half* w_hi;           // VQAE_PACK_SAME_MMA_F16 with premul: [11][C][C]
    const __half* w_lo;           // ... | VQAE_PACK_LO
    int n_tiles, H, W, tiles_x, tiles_per_img;
    FastDiv fd_tpi, fd_tx;
    float b1a, b1b, b2a, b2b, b3a, b3b, b4;
    float inv1, inv2, scale3;     // 1 / premul of W1, W2;  scale / premul of W3
};
struct TileS { int img, r0, c0; };

template <int C, bool K8>
struct WFragS {
    uint32_t b[C / 8][K8 ? 1 : C / 16][K8 ? 1 : 2];
    __device__ __forceinline__ void load_global(const __half* w, int g, int t) {
#pragma unroll
        for (int j = 0; j < C / 8; ++j)
#pragma unroll
            for (int s = 0; s < (K8 ? 1 : C / 16); ++s) {
                const __half* row = w + (8 * j + g) * C + 16 * s + 2 * t;
                b[j][s][0] = __ldg(reinterpret_cast<const uint32_t*>(row));
                if constexpr (!K8) b[j][s][1] = __ldg(reinterpret_cast<const uint32_t*>(row + 8));
            }
    }
};

// exact pre-activation on a scaled accumulator, split into fp16 hi / lo pairs
struct ActS {
    float pre, post, mul;
    __device__ __forceinline__ void operator()(float v0, float v1, uint32_t& hi, uint32_t& lo) const {
        const float f0 = elu1_tc(fmaf(v0, mul, pre)) + post, f1 = elu1_tc(fmaf(v1, mul, pre)) + post;
        const __half2 hh = __floats2half2_rn(f0, f1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
        hi = *reinterpret_cast<const uint32_t*>(&hh);
        lo = *reinterpret_cast<const uint32_t*>(&ll);
    }
};

template <int C, int TH>
__global__ void __launch_bounds__(SS_THREADS, SsCfg<C, TH>::MIN_CTAS)
same_block_mma_split_kernel(SsArgs a) {
    using Cfg = SsCfg<C, TH>;
    constexpr int KS = Cfg::KS, NT = Cfg::NT, UP = Cfg::UP, XP = Cfg::XP;
    constexpr bool K8 = Cfg::K8;
    constexpr uint32_t UL = Cfg::U_PLANE, WL = Cfg::W_SET;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = tc::smem_u32(smem);
    const uint32_t sU = sbase + Cfg::OFF_U;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    WFragS<C, K8> w1h, w1l, w3h, w3l;
    WFragS<C, K8> w2h[K8 ? 9 : 1], w2l[K8 ? 9 : 1];
    w1h.load_global(a.w_hi, g, t);
    w1l.load_global(a.w_lo, g, t);
    w3h.load_global(a.w_hi + 10 * C * C, g, t);
    w3l.load_global(a.w_lo + 10 * C * C, g, t);
    if constexpr (K8) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            w2h[tap].load_global(a.w_hi + (1 + tap) * C * C, g, t);
            w2l[tap].load_global(a.w_lo + (1 + tap) * C * C, g, t);
        }
    } else {
        for (int i = tid; i < 9 * C * C / 8; i += SS_THREADS) {
            const int row = i / (C / 8), piece = i % (C / 8);
            *reinterpret_cast<uint4*>(smem + Cfg::OFF_W + row * Cfg::WP + piece * 16) =
                __ldg(reinterpret_cast<const uint4*>(a.w_hi + C * C) + i);
            *reinterpret_cast<uint4*>(smem + Cfg::OFF_W + WL + row * Cfg::WP + piece * 16) =
                __ldg(reinterpret_cast<const uint4*>(a.w_lo + C * C) + i);
        }
    }
    const ActS act1{a.b1a, a.b1b, 1.f}, act2{a.b2a, a.b2b, a.inv1}, act3{a.b3a, a.b3b, a.inv2};

    auto decode = [&](int tile) {
        TileS tc_;
        if (a.fd_tpi.d == 1) { tc_.img = tile; tile = 0; }
        else { tc_.img = a.fd_tpi.div(tile); tile -= tc_.img * a.tiles_per_img; }
        const int ty = a.fd_tx.d == 1 ? tile : a.fd_tx.div(tile);
        tc_.r0 = ty * TH;
        tc_.c0 = (tile - ty * a.tiles_x) * SS_TW;
        return tc_;
    };
    const uint32_t bar0 = sbase + Cfg::OFF_BAR;
    if (tid == 0) {
        tc::mbar_init(bar0, 1);
        tc::mbar_init(bar0 + 8, 1);
        tc::fence_mbar_init();
    }
    __syncthreads();
    const int img_elems = a.H * a.W * C;
    auto issue_tile = [&](const TileS& tc_, int buf) {      // warps 0 and 1, all lanes (as in mma_same.cu)
        const float* ximg = a.x + (size_t)tc_.img * img_elems;
        const uint32_t dst0 = sbase + Cfg::OFF_X + buf * Cfg::X_BYTES;
        if (warp == 0) {
            const uint32_t bar = bar0 + 8 * buf;
            if (lane == 0) {
                tc::fence_proxy_async_smem();
                tc::mbar_arrive_expect_tx(bar, (TH + 2) * SS_TW * XP);
            }
            __syncwarp();
            if (lane < TH + 2) {
                int row = tc_.r0 - 1 + lane;
                row = row < 0 ? row + a.H : (row >= a.H ? row - a.H : row);
                tc::bulk_g2s(dst0 + (lane * SS_PW + 1) * XP, ximg + ((size_t)row * a.W + tc_.c0) * C,
                             SS_TW * XP, bar);
            }
        } else {
            constexpr int PPP = XP / 16;
            const int cl = tc_.c0 == 0 ? a.W - 1 : tc_.c0 - 1;
            const int cr = tc_.c0 + SS_TW == a.W ? 0 : tc_.c0 + SS_TW;
            for (int i = lane; i < 2 * (TH + 2) * PPP; i += 32) {
                const int piece = i % PPP, side = (i / PPP) & 1, lr = i / (2 * PPP);
                int row = tc_.r0 - 1 + lr;
                row = row < 0 ? row + a.H : (row >= a.H ? row - a.H : row);
                cp_async16(dst0 + (lr * SS_PW + (side ? SS_PW - 1 : 0)) * XP + piece * 16,
                           ximg + ((size_t)row * a.W + (side ? cr : cl)) * C + piece * 4);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };

    int tile = blockIdx.x;
    TileS cur = decode(tile < a.n_tiles ? tile : 0), nxt = cur;
    if (tile < a.n_tiles && warp < 2) issue_tile(cur, 0);
    int it = 0;
    for (; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        if (it > 0) cur = nxt;
        const int r0 = cur.r0, c0 = cur.c0;
        float* oimg = a.out + (size_t)cur.img * img_elems;
        const uint8_t* xs = smem + Cfg::OFF_X + buf * Cfg::X_BYTES;

        tc::fence_proxy_async_smem();
        if (warp == 1) asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        if (tile + (int)gridDim.x < a.n_tiles) {
            nxt = decode(tile + gridDim.x);
            if (warp < 2) issue_tile(nxt, buf ^ 1);
        }
        tc::mbar_wait(bar0 + 8 * buf, (it >> 1) & 1);

        // ================= stage 1 =================
        for (int m = warp; m < Cfg::MT1; m += SS_WARPS) {
            const int q0 = 16 * m + g, q1 = q0 + 8;
            uint32_t ah[KS][4], al[KS][4];
            if constexpr (K8) {
                const float2 v0 = *reinterpret_cast<const float2*>(xs + q0 * XP + 8 * t);
                const float2 v1 = *reinterpret_cast<const float2*>(xs + q1 * XP + 8 * t);
                act1(v0.x, v0.y, ah[0][0], al[0][0]);
                act1(v1.x, v1.y, ah[0][1], al[0][1]);
            } else {
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    const float4 v0 = *reinterpret_cast<const float4*>(xs + q0 * XP + 16 * t + 64 * s);
                    const float4 v1 = *reinterpret_cast<const float4*>(xs + q1 * XP + 16 * t + 64 * s);
                    act1(v0.x, v0.y, ah[s][0], al[s][0]);
                    act1(v1.x, v1.y, ah[s][1], al[s][1]);
                    act1(v0.z, v0.w, ah[s][2], al[s][2]);
                    act1(v1.z, v1.w, ah[s][3], al[s][3]);
                }
            }
            float d[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    if constexpr (K8) {
                        mma_1688(d[j], al[s][0], al[s][1], w1h.b[j][s][0]);
                        mma_1688(d[j], ah[s][0], ah[s][1], w1l.b[j][s][0]);
                        mma_1688(d[j], ah[s][0], ah[s][1], w1h.b[j][s][0]);
                    } else {
                        mma_16816(d[j], al[s], w1h.b[j][s][0], w1h.b[j][s][1]);
                        mma_16816(d[j], ah[s], w1l.b[j][s][0], w1l.b[j][s][1]);
                        mma_16816(d[j], ah[s], w1h.b[j][s][0], w1h.b[j][s][1]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                uint32_t h0, l0, h1, l1;
                act2(d[j][0], d[j][1], h0, l0);
                act2(d[j][2], d[j][3], h1, l1);
                const uint32_t o0 = Cfg::OFF_U + q0 * UP + (8 * j + 2 * t) * 2, o1 = Cfg::OFF_U + q1 * UP + (8 * j + 2 * t) * 2;
                *reinterpret_cast<uint32_t*>(smem + o0) = h0;
                *reinterpret_cast<uint32_t*>(smem + o0 + UL) = l0;
                *reinterpret_cast<uint32_t*>(smem + o1) = h1;
                *reinterpret_cast<uint32_t*>(smem + o1 + UL) = l1;
            }
        }
        __syncthreads();

        // ================= stages 2 + 3 =================
        for (int mt = warp; mt < Cfg::MT2; mt += SS_WARPS) {
            const int r = mt >> 1, cb = (mt & 1) * 16;
            const int qc = (r + 1) * SS_PW + cb + 1;
            float d[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
            const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;
            const uint32_t lbase = sU + (uint32_t)(qc + lrow) * UP + (K8 ? 0 : (lane >> 4) * 16);
            const uint32_t wbase = sbase + Cfg::OFF_W + (uint32_t)((lane & 7) + (lane >> 4) * 8) * Cfg::WP +
                                   ((lane >> 3) & 1) * 16;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int shift = (tap / 3 - 1) * SS_PW + (tap % 3 - 1);
                if constexpr (K8) {
                    uint32_t h0, h1, l0, l1;
                    ldmatrix_x2(h0, h1, lbase + shift * UP);
                    ldmatrix_x2(l0, l1, lbase + UL + shift * UP);
                    mma_1688(d[0], l0, l1, w2h[tap].b[0][0][0]);
                    mma_1688(d[0], h0, h1, w2l[tap].b[0][0][0]);
                    mma_1688(d[0], h0, h1, w2h[tap].b[0][0][0]);
                } else {
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        uint32_t ah[4], al[4];
                        ldmatrix_x4(ah, lbase + shift * UP + s * 32);
                        ldmatrix_x4(al, lbase + UL + shift * UP + s * 32);
#pragma unroll
                        for (int p = 0; p < NT / 2; ++p) {
                            uint32_t bh[4], bl[4];
                            ldmatrix_x4(bh, wbase + (uint32_t)(tap * C + 16 * p) * Cfg::WP + s * 32);
                            ldmatrix_x4(bl, wbase + WL + (uint32_t)(tap * C + 16 * p) * Cfg::WP + s * 32);
                            mma_16816(d[2 * p], al, bh[0], bh[1]);
                            mma_16816(d[2 * p + 1], al, bh[2], bh[3]);
                            mma_16816(d[2 * p], ah, bl[0], bl[1]);
                            mma_16816(d[2 * p + 1], ah, bl[2], bl[3]);
                            mma_16816(d[2 * p], ah, bh[0], bh[1]);
                            mma_16816(d[2 * p + 1], ah, bh[2], bh[3]);
                        }
                    }
                }
            }
            uint32_t vh[KS][4], vl[KS][4];
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                if constexpr (K8) {
                    act3(d[0][0], d[0][1], vh[s][0], vl[s][0]);
                    act3(d[0][2], d[0][3], vh[s][1], vl[s][1]);
                } else {
                    act3(d[2 * s][0], d[2 * s][1], vh[s][0], vl[s][0]);
                    act3(d[2 * s][2], d[2 * s][3], vh[s][1], vl[s][1]);
                    act3(d[2 * s + 1][0], d[2 * s + 1][1], vh[s][2], vl[s][2]);
                    act3(d[2 * s + 1][2], d[2 * s + 1][3], vh[s][3], vl[s][3]);
                }
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    if constexpr (K8) {
                        mma_1688(d[j], vl[s][0], vl[s][1], w3h.b[j][s][0]);
                        mma_1688(d[j], vh[s][0], vh[s][1], w3l.b[j][s][0]);
                        mma_1688(d[j], vh[s][0], vh[s][1], w3h.b[j][s][0]);
                    } else {
                        mma_16816(d[j], vl[s], w3h.b[j][s][0], w3h.b[j][s][1]);
                        mma_16816(d[j], vh[s], w3l.b[j][s][0], w3l.b[j][s][1]);
                        mma_16816(d[j], vh[s], w3h.b[j][s][0], w3h.b[j][s][1]);
                    }
                }
            }
            const int qa = qc + g, qb = qa + 8;
            const size_t ooff0 = ((size_t)(r0 + r) * a.W + c0 + cb + g) * C + (K8 ? 2 : 4) * t;
            const size_t ooff1 = ooff0 + (size_t)8 * C;
            if constexpr (K8) {
                const float2 x0 = *reinterpret_cast<const float2*>(xs + qa * XP + 8 * t);
                const float2 x1 = *reinterpret_cast<const float2*>(xs + qb * XP + 8 * t);
                float2 o0, o1;
                o0.x = fmaf(d[0][0], a.scale3, a.b4) + x0.x;
                o0.y = fmaf(d[0][1], a.scale3, a.b4) + x0.y;
                o1.x = fmaf(d[0][2], a.scale3, a.b4) + x1.x;
                o1.y = fmaf(d[0][3], a.scale3, a.b4) + x1.y;
                *reinterpret_cast<float2*>(oimg + ooff0) = o0;
                *reinterpret_cast<float2*>(oimg + ooff1) = o1;
            } else {
#pragma unroll
                for (int p = 0; p < NT / 2; ++p) {
                    const int oa = 16 * t + 64 * p;
                    const float4 x0 = *reinterpret_cast<const float4*>(xs + qa * XP + oa);
                    const float4 x1 = *reinterpret_cast<const float4*>(xs + qb * XP + oa);
                    const float (&e)[4] = d[2 * p], (&f)[4] = d[2 * p + 1];
                    float4 o0, o1;
                    o0.x = fmaf(e[0], a.scale3, a.b4) + x0.x;
                    o0.y = fmaf(e[1], a.scale3, a.b4) + x0.y;
                    o0.z = fmaf(f[0], a.scale3, a.b4) + x0.z;
                    o0.w = fmaf(f[1], a.scale3, a.b4) + x0.w;
                    o1.x = fmaf(e[2], a.scale3, a.b4) + x1.x;
                    o1.y = fmaf(e[3], a.scale3, a.b4) + x1.y;
                    o1.z = fmaf(f[2], a.scale3, a.b4) + x1.z;
                    o1.w = fmaf(f[3], a.scale3, a.b4) + x1.w;
                    *reinterpret_cast<float4*>(oimg + ooff0 + 16 * p) = o0;
                    *reinterpret_cast<float4*>(oimg + ooff1 + 16 * p) = o1;
                }
            }
        }
    }
}

template <int C, int TH>
int launch_same_split(SsArgs a, int64_t B, int sm_count, cudaStream_t stream) {
    using Cfg = SsCfg<C, TH>;
    auto kern = same_block_mma_split_kernel<C, TH>;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set.cur() = true;
    }
    a.tiles_x = a.W / SS_TW;
    a.tiles_per_img = (a.H / TH) * a.tiles_x;
    const int64_t nt = B * a.tiles_per_img;
    if (nt > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    a.n_tiles = (int)nt;
    a.fd_tpi = make_fastdiv(a.tiles_per_img);
    a.fd_tx = make_fastdiv(a.tiles_x);
    const int cap = sm_count * Cfg::MIN_CTAS;
    const int grid = a.n_tiles < cap ? a.n_tiles : cap;
    kern<<<grid, SS_THREADS, Cfg::SMEM, stream>>>(a);
    return check_launch();
}

}  // namespace

bool same_block_mma_split_supported(int H, int W, int C) {
    return (C == 8 || C == 16) && H >= 8 && W >= SS_TW && H % 8 == 0 && W % SS_TW == 0;
}

int same_block_mma_split(const float* x, float* out, const void* w_hi, const void* w_lo,
                         const float* scalars8, const float* premul3, int64_t B, int H, int W, int C,
                         int sm_count, cudaStream_t stream) {
    if (!x || !out || !w_hi || !w_lo || !scalars8 || !premul3 || B <= 0) return VQAE_ERR_BAD_ARG;
    if (x == out) return VQAE_ERR_BAD_ARG;
    if (!same_block_mma_split_supported(H, W, C)) return VQAE_ERR_UNSUPPORTED;
    for (int i = 0; i < 3; ++i)
        if (!(premul3[i] > 0.f)) return VQAE_ERR_BAD_ARG;
    SsArgs a;
    a.x = x; a.out = out;
    a.w_hi = reinterpret_cast<const __half*>(w_hi);
    a.w_lo = reinterpret_cast<const __half*>(w_lo);
    a.H = H; a.W = W;
    a.b1a = scalars8[0]; a.b1b = scalars8[1]; a.b2a = scalars8[2]; a.b2b = scalars8[3];
    a.b3a = scalars8[4]; a.b3b = scalars8[5]; a.b4 = scalars8[6];
    a.inv1 = 1.f / premul3[0]; a.inv2 = 1.f / premul3[1]; a.scale3 = scalars8[7] / premul3[2];
    if (C == 8) return launch_same_split<8, 8>(a, B, sm_count, stream);
    return launch_same_split<16, 8>(a, B, sm_count, stream);
}

}  // namespace vqae
