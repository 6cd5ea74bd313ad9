// PreActFixupResBlock 'same' at LOW channel counts (C = 8, 16, 32 -- the full-resolution levels of
// the encoder / decoder pyramids, layers/conv_block.py:196-216 at 256^2 / 128^2 / 64^2) as ONE kernel
// built on warp-level tensor-core MMAs (mma.sync m16n8k16 / m16n8k8, fp16 operands, fp32
// accumulation) with the three GEMMs chained through REGISTERS.
//
// Why not tcgen05 here: with N = C <= 32 a tcgen05.mma moves a 128 x 16 operand slab through shared
// memory for 1/16 .. 1/4 of the work of a full-rate instruction, every stage boundary is a
// TMEM -> register -> shared-memory round trip behind a CTA-wide barrier, and one CTA holds 512 pixels
// in lock step.  Measured on B200: the tcgen05 tile kernel spends ~31 k cycles
// per 512-pixel tile at ~25 % issue utilisation and does not speed up when its HBM bytes are halved
// -- it is latency bound, not byte bound.  These layers hold 26 % of the encoder's FLOPs, but their cost
// is the activation arithmetic (three ELUs per channel and pixel: half of all instructions) and the
// memory stream, so the design that fits is many independent warps, each carrying 16 pixels through
// the whole block, fed by an asynchronous double-buffered tile load:
//
//   bulk copy  x tile + 1-pixel circular halo (fp32, three UBLKCP per tile row with wrapped sources,
//              completion on an mbarrier) -> shared memory, issued one tile ahead by one warp;
//              nothing in the block waits on global-memory latency
//   stage 1    (pointwise, on all (TH + 2) x 34 halo'd pixels, 16 per warp step)
//       A1 = f16(elu(x + b1a) + b1b)        built straight in the A-fragment registers
//       D1 = A1 . W1^T                      mma.sync, W1 fragments live in registers
//       U  = f16(elu(D1 + b2a) + b2b)       -> shared memory [pixel][C] (pitch C*2 + 16 B)
//   __syncthreads
//   stages 2 + 3 (interior, 16 consecutive pixels of a row per warp step)
//       D2 = sum_taps U[q + dy*34 + dx] . W2[tap]^T     A (and W2) fragments by ldmatrix
//       V  = f16(elu(D2 + b3a) + b3b)       the accumulator fragment IS the next A fragment
//       D3 = V . W3^T
//       out = x + scale * D3 + b4           fp32, residual from the staged x tile, 128-bit stores
//
// Two CTA barriers per tile, no tensor memory, no asynchronous-proxy fences.
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "mma_common.cuh"
#include "tc_common.cuh"

namespace vqae {
namespace {

using namespace mma;

constexpr int MS_TW = 32, MS_PW = MS_TW + 2;
constexpr int MS_WARPS = 8;
constexpr int MS_THREADS = MS_WARPS * 32;

template <int C, int TH>
struct MsCfg {
    static constexpr bool K8 = (C == 8);                 // m16n8k8 (no zero padding of K)
    static constexpr int KS = K8 ? 1 : C / 16;           // k-steps per GEMM
    static constexpr int NT = C / 8;                     // 8-wide n-tiles
    static constexpr int NPAD = (TH + 2) * MS_PW;        // halo'd pixels in padded-linear order
    static constexpr int MT1 = (NPAD + 15) / 16;         // M-tiles in stage 1
    static constexpr int MT2 = TH * MS_TW / 16;          // M-tiles in stages 2 + 3
    static constexpr int UP = K8 ? 16 : C * 2 + 16;      // bytes per pixel of U: 16 * odd -> ldmatrix
                                                         // and the fragment stores are conflict-free
    static constexpr int XP = C * 4;                     // bytes per pixel of the staged x tile
    static constexpr int WP = C * 2 + 16;                // W2 row pitch in shared memory
    static constexpr bool W2_SMEM = !K8;                 // C = 8: all weights in registers
    static constexpr uint32_t X_BYTES = (uint32_t)(MT1 * 16) * XP;          // incl. slack rows
    static constexpr uint32_t U_BYTES = (uint32_t)(MT1 * 16) * UP;
    static constexpr int NXB = (C == 32) ? 1 : 2;       // x buffers (C = 32: two would leave one CTA per SM)
    static constexpr uint32_t OFF_X = 0;
    static constexpr uint32_t OFF_U = OFF_X + NXB * X_BYTES;
    static constexpr uint32_t OFF_W = OFF_U + U_BYTES;
    static constexpr uint32_t W_BYTES = W2_SMEM ? 9 * C * WP : 0;
    static constexpr uint32_t OFF_BAR = OFF_W + W_BYTES;                    // one mbarrier per x buffer
    static constexpr uint32_t SMEM = OFF_BAR + 16;
    static constexpr int FIT = (227 * 1024) / (int)(SMEM + 1024);          // CTAs per SM by shared memory
    static constexpr int MIN_CTAS = FIT >= 4 ? 4 : (FIT >= 3 ? 3 : (FIT >= 2 ? 2 : 1));
};

struct MsArgs {
    const float* x;               // NHWC fp32 [B,H,W,C]
    float* out;                   // NHWC fp32 [B,H,W,C]
    const __half* w;              // [11][C][C]: W1 | W2 tap 0..8 | W3, each [out ch][in ch]
    int n_tiles, H, W, tiles_x, tiles_per_img;
    FastDiv fd_tpi, fd_tx;
    float b1a, b1b, b2a, b2b, b3a, b3b, b4, scale;
};
struct TileC { int img, r0, c0; };

// B fragments of one C x C matrix W[n][k] (row-major, fp16) from global memory: for n-tile j, k-step s
//   b0 = W[8j + g][16s + 2t, +1],  b1 = W[8j + g][16s + 2t + 8, +9]      (K8: b0 only, k = 2t, 2t + 1)
template <int C, bool K8>
struct WFrag {
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

template <int C, int TH>
__global__ void __launch_bounds__(MS_THREADS, MsCfg<C, TH>::MIN_CTAS)
same_block_mma_kernel(MsArgs a) {
    using Cfg = MsCfg<C, TH>;
    constexpr int KS = Cfg::KS, NT = Cfg::NT, UP = Cfg::UP, XP = Cfg::XP, NPAD = Cfg::NPAD;
    constexpr bool K8 = Cfg::K8;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = tc::smem_u32(smem);
    const uint32_t sU = sbase + Cfg::OFF_U;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    // ---- weights: W1 / W3 fragments in registers; W2 in registers (C = 8) or shared memory ----
    WFrag<C, K8> w1, w3;
    WFrag<C, K8> w2r[K8 ? 9 : 1];
    w1.load_global(a.w, g, t);
    w3.load_global(a.w + 10 * C * C, g, t);
    if constexpr (K8) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) w2r[tap].load_global(a.w + (1 + tap) * C * C, g, t);
    } else {
        for (int i = tid; i < 9 * C * C / 8; i += MS_THREADS) {          // 16-byte pieces of W2
            const int row = i / (C / 8), piece = i % (C / 8);
            *reinterpret_cast<uint4*>(smem + Cfg::OFF_W + row * Cfg::WP + piece * 16) =
                __ldg(reinterpret_cast<const uint4*>(a.w + C * C) + i);
        }
    }
    const ActC act1(a.b1a, a.b1b), act2(a.b2a, a.b2b), act3(a.b3a, a.b3b);

    auto decode = [&](int tile) {
        TileC tc_;
        if (a.fd_tpi.d == 1) { tc_.img = tile; tile = 0; }
        else { tc_.img = a.fd_tpi.div(tile); tile -= tc_.img * a.tiles_per_img; }
        const int ty = a.fd_tx.d == 1 ? tile : a.fd_tx.div(tile);
        tc_.r0 = ty * TH;
        tc_.c0 = (tile - ty * a.tiles_x) * MS_TW;
        return tc_;
    };

    // ---- asynchronous load of one halo'd x tile into x buffer `buf`: bulk copies (UBLKCP) issued by
    //      warp 0, three per tile row (left halo pixel | 32 pixels | right halo pixel, each with its
    //      wrapped source), completion counted in bytes on the buffer's mbarrier ----
    const uint32_t bar0 = sbase + Cfg::OFF_BAR;
    if (tid == 0) {
        tc::mbar_init(bar0, 1);
        tc::mbar_init(bar0 + 8, 1);
        tc::fence_mbar_init();
    }
    __syncthreads();
    const int img_elems = a.H * a.W * C;
    auto issue_tile = [&](const TileC& tc_, int buf) {      // call from warps 0 and 1, all lanes
        const float* ximg = a.x + (size_t)tc_.img * img_elems;
        const uint32_t dst0 = sbase + Cfg::OFF_X + buf * Cfg::X_BYTES;
        if (warp == 0) {
            // the 32 interior columns of every row: one bulk copy (UBLKCP) per row
            const uint32_t bar = bar0 + 8 * buf;
            if (lane == 0) {
                tc::fence_proxy_async_smem();               // earlier generic reads of this buffer
                tc::mbar_arrive_expect_tx(bar, (TH + 2) * MS_TW * XP);
            }
            __syncwarp();
            if (lane < TH + 2) {
                int row = tc_.r0 - 1 + lane;
                row = row < 0 ? row + a.H : (row >= a.H ? row - a.H : row);
                tc::bulk_g2s(dst0 + (lane * MS_PW + 1) * XP, ximg + ((size_t)row * a.W + tc_.c0) * C,
                             MS_TW * XP, bar);
            }
        } else {
            // the two halo columns (wrapped): 16-byte cp.async pieces, waited for by this warp at
            // the top of the next iteration
            constexpr int PPP = XP / 16;
            const int cl = tc_.c0 == 0 ? a.W - 1 : tc_.c0 - 1;
            const int cr = tc_.c0 + MS_TW == a.W ? 0 : tc_.c0 + MS_TW;
            for (int i = lane; i < 2 * (TH + 2) * PPP; i += 32) {
                const int piece = i % PPP, side = (i / PPP) & 1, lr = i / (2 * PPP);
                int row = tc_.r0 - 1 + lr;
                row = row < 0 ? row + a.H : (row >= a.H ? row - a.H : row);
                cp_async16(dst0 + (lr * MS_PW + (side ? MS_PW - 1 : 0)) * XP + piece * 16,
                           ximg + ((size_t)row * a.W + (side ? cr : cl)) * C + piece * 4);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };

    int tile = blockIdx.x;
    TileC cur = decode(tile < a.n_tiles ? tile : 0), nxt = cur;
    if (Cfg::NXB == 2 && tile < a.n_tiles && warp < 2) issue_tile(cur, 0);
    int it = 0;
    for (; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int buf = Cfg::NXB == 2 ? (it & 1) : 0;
        if (it > 0) cur = Cfg::NXB == 2 ? nxt : decode(tile);
        const int r0 = cur.r0, c0 = cur.c0;
        float* oimg = a.out + (size_t)cur.img * img_elems;
        const uint8_t* xs = smem + Cfg::OFF_X + buf * Cfg::X_BYTES;

        // the previous tile is finished: U and the other x buffer are free; every thread's generic
        // reads of that buffer are ordered before the asynchronous-proxy refill
        tc::fence_proxy_async_smem();
        if (Cfg::NXB == 2 && warp == 1) asm volatile("cp.async.wait_all;" ::: "memory");   // halo columns of `buf`
        __syncthreads();
        if constexpr (Cfg::NXB == 1) {
            if (warp < 2) issue_tile(cur, 0);
            if (warp == 1) asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
        } else if (tile + (int)gridDim.x < a.n_tiles) {
            nxt = decode(tile + gridDim.x);
            if (warp < 2) issue_tile(nxt, buf ^ 1);
        }
        tc::mbar_wait(bar0 + 8 * buf, (Cfg::NXB == 2 ? (it >> 1) : it) & 1);     // x tile `buf` landed

        // Both stages carry TWO M-tiles per warp step where the tile count allows: the per-M-tile work
        // is a chain of dependent MMAs / MUFUs, so a second independent chain doubles what a warp
        // keeps in flight, and in stage 2 the W2 fragments are read from shared memory once per pair.
        // ================= stage 1: U = f16(elu(W1 . f16(elu(x + b1a) + b1b) + b2a) + b2b) =============
        auto stage1 = [&](auto np_c, int m0) {
            constexpr int NP = decltype(np_c)::value;
            uint32_t af[NP][KS][4];
#pragma unroll
            for (int u = 0; u < NP; ++u) {
                const int q0 = 16 * (m0 + u * MS_WARPS) + g, q1 = q0 + 8;   // rows g, g + 8 (slack rows allocated)
                if constexpr (K8) {
                    const float2 v0 = *reinterpret_cast<const float2*>(xs + q0 * XP + 8 * t);
                    const float2 v1 = *reinterpret_cast<const float2*>(xs + q1 * XP + 8 * t);
                    af[u][0][0] = act1(v0.x, v0.y);
                    af[u][0][1] = act1(v1.x, v1.y);
                } else {
                    // lane (g, t) takes channels 16s + 4t .. 4t + 3: fragment slots k' = 2t, 2t + 1,
                    // 2t + 8, 2t + 9 -- the k order of W1 is permuted to match at pack time (pack.cu)
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        const float4 v0 = *reinterpret_cast<const float4*>(xs + q0 * XP + 16 * t + 64 * s);
                        const float4 v1 = *reinterpret_cast<const float4*>(xs + q1 * XP + 16 * t + 64 * s);
                        af[u][s][0] = act1(v0.x, v0.y);
                        af[u][s][1] = act1(v1.x, v1.y);
                        af[u][s][2] = act1(v0.z, v0.w);
                        af[u][s][3] = act1(v1.z, v1.w);
                    }
                }
            }
            float d[NP][NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
#pragma unroll
                for (int u = 0; u < NP; ++u) d[u][j][0] = d[u][j][1] = d[u][j][2] = d[u][j][3] = 0.f;
#pragma unroll
                for (int s = 0; s < KS; ++s)
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        if constexpr (K8) mma_1688(d[u][j], af[u][s][0], af[u][s][1], w1.b[j][s][0]);
                        else mma_16816(d[u][j], af[u][s], w1.b[j][s][0], w1.b[j][s][1]);
                    }
            }
#pragma unroll
            for (int u = 0; u < NP; ++u) {
                const int q0 = 16 * (m0 + u * MS_WARPS) + g, q1 = q0 + 8;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    *reinterpret_cast<uint32_t*>(smem + Cfg::OFF_U + q0 * UP + (8 * j + 2 * t) * 2) = act2(d[u][j][0], d[u][j][1]);
                    *reinterpret_cast<uint32_t*>(smem + Cfg::OFF_U + q1 * UP + (8 * j + 2 * t) * 2) = act2(d[u][j][2], d[u][j][3]);
                }
            }
        };
        for (int m = warp; m < Cfg::MT1; m += 2 * MS_WARPS) {
            if (m + MS_WARPS < Cfg::MT1) stage1(std::integral_constant<int, 2>{}, m);
            else stage1(std::integral_constant<int, 1>{}, m);
        }
        __syncthreads();

        // ================= stages 2 + 3 =================
        auto stage23 = [&](auto np_c, int mt0) {
            constexpr int NP = decltype(np_c)::value;
            int qc[NP], rr[NP], cbb[NP];
            uint32_t lbase[NP];
            // ldmatrix row address of this lane: matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15)
            const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
            for (int u = 0; u < NP; ++u) {
                const int mt = mt0 + u * MS_WARPS;
                rr[u] = mt >> 1;                                   // tile row
                cbb[u] = (mt & 1) * 16;                            // first column
                qc[u] = (rr[u] + 1) * MS_PW + cbb[u] + 1;          // padded-linear index of pixel 0
                lbase[u] = sU + (uint32_t)(qc[u] + lrow) * UP + (K8 ? 0 : (lane >> 4) * 16);
            }
            float d[NP][NT][4];
#pragma unroll
            for (int u = 0; u < NP; ++u)
#pragma unroll
                for (int j = 0; j < NT; ++j) d[u][j][0] = d[u][j][1] = d[u][j][2] = d[u][j][3] = 0.f;
            // W2 fragments: matrices (n 0-7 | 8-15) x (k 0-7 | 8-15) of an n-tile pair
            const uint32_t wbase = sbase + Cfg::OFF_W + (uint32_t)((lane & 7) + (lane >> 4) * 8) * Cfg::WP +
                                   ((lane >> 3) & 1) * 16;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int shift = (tap / 3 - 1) * MS_PW + (tap % 3 - 1);
                if constexpr (K8) {
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        uint32_t a0, a1;
                        ldmatrix_x2(a0, a1, lbase[u] + shift * UP);
                        mma_1688(d[u][0], a0, a1, w2r[tap].b[0][0][0]);
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        uint32_t af[NP][4];
#pragma unroll
                        for (int u = 0; u < NP; ++u) ldmatrix_x4(af[u], lbase[u] + shift * UP + s * 32);
#pragma unroll
                        for (int p = 0; p < NT / 2; ++p) {
                            uint32_t bf[4];          // b(2p, s)[0], b(2p, s)[1], b(2p+1, s)[0], b(2p+1, s)[1]
                            ldmatrix_x4(bf, wbase + (uint32_t)(tap * C + 16 * p) * Cfg::WP + s * 32);
#pragma unroll
                            for (int u = 0; u < NP; ++u) {
                                mma_16816(d[u][2 * p], af[u], bf[0], bf[1]);
                                mma_16816(d[u][2 * p + 1], af[u], bf[2], bf[3]);
                            }
                        }
                    }
                }
            }
            // V = f16(elu(D2 + b3a) + b3b): accumulator fragments of n-tiles (2s, 2s + 1) are the A
            // fragment of k-step s
            uint32_t vf[NP][KS][4];
#pragma unroll
            for (int u = 0; u < NP; ++u)
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    if constexpr (K8) {
                        vf[u][s][0] = act3(d[u][0][0], d[u][0][1]);
                        vf[u][s][1] = act3(d[u][0][2], d[u][0][3]);
                    } else {
                        vf[u][s][0] = act3(d[u][2 * s][0], d[u][2 * s][1]);
                        vf[u][s][1] = act3(d[u][2 * s][2], d[u][2 * s][3]);
                        vf[u][s][2] = act3(d[u][2 * s + 1][0], d[u][2 * s + 1][1]);
                        vf[u][s][3] = act3(d[u][2 * s + 1][2], d[u][2 * s + 1][3]);
                    }
                }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
#pragma unroll
                for (int u = 0; u < NP; ++u) d[u][j][0] = d[u][j][1] = d[u][j][2] = d[u][j][3] = 0.f;
#pragma unroll
                for (int s = 0; s < KS; ++s)
#pragma unroll
                    for (int u = 0; u < NP; ++u) {
                        if constexpr (K8) mma_1688(d[u][j], vf[u][s][0], vf[u][s][1], w3.b[j][s][0]);
                        else mma_16816(d[u][j], vf[u][s], w3.b[j][s][0], w3.b[j][s][1]);
                    }
            }
            // out = x + scale * D3 + b4: residual from the staged tile; C >= 16: the output-channel
            // order of W3 is permuted at pack time so that lane (g, t) owns channels 16p + 4t .. + 3
#pragma unroll
            for (int u = 0; u < NP; ++u) {
                const int qa = qc[u] + g, qb = qa + 8;
                const size_t ooff0 = ((size_t)(r0 + rr[u]) * a.W + c0 + cbb[u] + g) * C + (K8 ? 2 : 4) * t;
                const size_t ooff1 = ooff0 + (size_t)8 * C;
                if constexpr (K8) {
                    const float2 x0 = *reinterpret_cast<const float2*>(xs + qa * XP + 8 * t);
                    const float2 x1 = *reinterpret_cast<const float2*>(xs + qb * XP + 8 * t);
                    float2 o0, o1;
                    o0.x = fmaf(d[u][0][0], a.scale, a.b4) + x0.x;
                    o0.y = fmaf(d[u][0][1], a.scale, a.b4) + x0.y;
                    o1.x = fmaf(d[u][0][2], a.scale, a.b4) + x1.x;
                    o1.y = fmaf(d[u][0][3], a.scale, a.b4) + x1.y;
                    *reinterpret_cast<float2*>(oimg + ooff0) = o0;
                    *reinterpret_cast<float2*>(oimg + ooff1) = o1;
                } else {
#pragma unroll
                    for (int p = 0; p < NT / 2; ++p) {
                        const int oa = 16 * t + 64 * p;
                        const float4 x0 = *reinterpret_cast<const float4*>(xs + qa * XP + oa);
                        const float4 x1 = *reinterpret_cast<const float4*>(xs + qb * XP + oa);
                        const float (&e)[4] = d[u][2 * p], (&f)[4] = d[u][2 * p + 1];
                        float4 o0, o1;
                        o0.x = fmaf(e[0], a.scale, a.b4) + x0.x;
                        o0.y = fmaf(e[1], a.scale, a.b4) + x0.y;
                        o0.z = fmaf(f[0], a.scale, a.b4) + x0.z;
                        o0.w = fmaf(f[1], a.scale, a.b4) + x0.w;
                        o1.x = fmaf(e[2], a.scale, a.b4) + x1.x;
                        o1.y = fmaf(e[3], a.scale, a.b4) + x1.y;
                        o1.z = fmaf(f[2], a.scale, a.b4) + x1.z;
                        o1.w = fmaf(f[3], a.scale, a.b4) + x1.w;
                        *reinterpret_cast<float4*>(oimg + ooff0 + 16 * p) = o0;
                        *reinterpret_cast<float4*>(oimg + ooff1 + 16 * p) = o1;
                    }
                }
            }
        };
        constexpr int NP23 = (C <= 16) ? 2 : 1;          // C = 32: two chains do not fit the register budget
        static_assert(Cfg::MT2 % (MS_WARPS * NP23) == 0, "M-tile pairs tile the stage");
        for (int mt = warp; mt < Cfg::MT2; mt += NP23 * MS_WARPS)
            stage23(std::integral_constant<int, NP23>{}, mt);
    }
}

template <int C, int TH>
int launch_same_mma(MsArgs a, int64_t B, int sm_count, cudaStream_t stream) {
    using Cfg = MsCfg<C, TH>;
    auto kern = same_block_mma_kernel<C, TH>;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)Cfg::SMEM));
        attr_set.cur() = true;
    }
    a.tiles_x = a.W / MS_TW;
    a.tiles_per_img = (a.H / TH) * a.tiles_x;
    const int64_t nt = B * a.tiles_per_img;
    if (nt > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    a.n_tiles = (int)nt;
    a.fd_tpi = make_fastdiv(a.tiles_per_img);
    a.fd_tx = make_fastdiv(a.tiles_x);
    const int cap = sm_count * Cfg::MIN_CTAS;
    const int grid = a.n_tiles < cap ? a.n_tiles : cap;
    kern<<<grid, MS_THREADS, Cfg::SMEM, stream>>>(a);
    return check_launch();
}

}  // namespace

bool same_block_mma_supported(int H, int W, int C) {
    return (C == 8 || C == 16 || C == 32) && H >= 8 && W >= MS_TW && H % 8 == 0 && W % MS_TW == 0;
}

int same_block_mma(const float* x, float* out, const void* w_packed, const float* scalars8,
                   int64_t B, int H, int W, int C, int sm_count, cudaStream_t stream) {
    if (!x || !out || !w_packed || !scalars8 || B <= 0) return VQAE_ERR_BAD_ARG;
    if (x == out) return VQAE_ERR_BAD_ARG;
    if (!same_block_mma_supported(H, W, C)) return VQAE_ERR_UNSUPPORTED;
    MsArgs a;
    a.x = x; a.out = out; a.w = reinterpret_cast<const __half*>(w_packed);
    a.H = H; a.W = W;
    a.b1a = scalars8[0]; a.b1b = scalars8[1]; a.b2a = scalars8[2]; a.b2b = scalars8[3];
    a.b3a = scalars8[4]; a.b3b = scalars8[5]; a.b4 = scalars8[6]; a.scale = scalars8[7];
    switch (C) {
        case 8: return H % 16 == 0 ? launch_same_mma<8, 16>(a, B, sm_count, stream)
                                   : launch_same_mma<8, 8>(a, B, sm_count, stream);
        case 16: return launch_same_mma<16, 8>(a, B, sm_count, stream);
        case 32: return launch_same_mma<32, 8>(a, B, sm_count, stream);
    }
    return VQAE_ERR_UNSUPPORTED;
}

}  // namespace vqae
