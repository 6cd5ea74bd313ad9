// The FRONT END of the encoder as ONE kernel: input normalisation + in_stem (3x3 zero-pad conv 3 -> 8,
// vq_ae/model.py:141,198) + the C = 8 'same' block at full resolution + the 'down' block 8 -> 16
// (the first DownBlock, model.py:144-148 / layers/conv_block.py:94-129,196-216).  Unfused these are
// three launches that write and re-read two 8-channel full-resolution fp32 tensors (2 x 537 MB at
// batch 256 of 256^2 patches); fused, the kernel reads the uint8 image (50 MB) and writes the
// 16-channel half-resolution tensor (268 MB), everything in between lives in shared memory and
// registers of the CTA that owns a 16 x 32 pixel tile:
//
//   stage   input window (tile + 2-pixel apron, zero outside the image = the stem's padding),
//           normalised, split hi + lo into fp16, one pixel = 4 halves (R, G, B, 0)
//   stem    x0 = in_stem(window) on the tile + its 1-pixel ring, on tensor cores: per kernel row ky
//           one m16n8k16 MMA whose K = 4 pixels x 4 halves; lane t of a fragment row owns window
//           pixel kx = t (weights zero on the 4th pixel and the 4th channel), so its hi and lo
//           operand registers are ONE 128-bit load of the staged pixel; split operands (hi.hi +
//           lo.hi + hi.lo) keep it fp32-accurate.  The ring of the C = 8 block wraps around the IMAGE
//           (circular padding), so ring pixels of tiles on the image border are evaluated separately
//           at their wrapped coordinates (plain fp32, from global memory).
//   same    the three chained GEMMs of the C = 8 block exactly as in mma_same.cu (K = 8 MMAs, weights
//           in registers, U through shared memory), result y kept in shared memory
//   down    the 'down' block exactly as in mma_down.cu (register-chained GEMMs over the four 2x2
//           positions), input from y, output to global memory
//
// Operand precision of the two blocks = the "fp16" path (fp16 operands, fp32 accumulation, fast
// ELU): the kernel replaces stem_in + same_block_mma + down_block_mma of that path.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"
#include "mma_common.cuh"
#include "tc_common.cuh"

namespace vqae {
namespace {

using namespace mma;

constexpr int FF_TH = 16, FF_TW = 32, FF_PW = FF_TW + 2;
constexpr int FF_NPAD = (FF_TH + 2) * FF_PW;          // 612 pixels of tile + ring, padded-linear
constexpr int FF_MT1 = (FF_NPAD + 15) / 16;            // 39 M-tiles (624 rows incl. slack)
constexpr int FF_MT2 = FF_TH * FF_TW / 16;             // 32 interior M-tiles
constexpr int FF_SROWS = FF_TH + 5;                    // staged rows: image rows r0-2 .. r0+TH+1, + slack
constexpr int FF_SPW = FF_TW + 6;                      // staged pixels per row (36 real + 2 zero)
constexpr int FF_SREAL_R = FF_TH + 4, FF_SREAL_C = FF_TW + 4;
constexpr int FF_THREADS = 256, FF_WARPS = 8;

// down-block weights in shared memory: the layout of MdCfg<8> in mma_down.cu
constexpr int FD_CI = 8, FD_CO = 16, FD_WPI = 16, FD_WPO = FD_CO * 2 + 16;
constexpr uint32_t FD_OFF_W1 = 0;
constexpr uint32_t FD_OFF_W2 = FD_OFF_W1 + FD_CO * FD_WPI;
constexpr uint32_t FD_OFF_W3 = FD_OFF_W2 + 4 * FD_CO * FD_WPO;
constexpr uint32_t FD_OFF_WS = FD_OFF_W3 + FD_CO * FD_WPO;
constexpr uint32_t FD_BYTES = FD_OFF_WS + 4 * FD_CO * FD_WPI;

constexpr uint32_t S_BYTES = FF_SROWS * FF_SPW * 16;                     // staged pixel = 16 bytes:
                                                                         // hi (R, G, B, 0) | lo (R, G, B, 0)
constexpr uint32_t Y_BYTES = FF_TH * FF_TW * 32;                         // y: [pixel][8] fp32
constexpr uint32_t OFF_S = 0;                                            // y aliases it later
constexpr uint32_t OFF_XS = (S_BYTES > Y_BYTES ? S_BYTES : Y_BYTES);
constexpr uint32_t OFF_U = OFF_XS + FF_MT1 * 16 * 32;                    // x0: [q][8] fp32
constexpr uint32_t OFF_DW = OFF_U + FF_MT1 * 16 * 16;                    // U:  [q][8] fp16
constexpr uint32_t FF_SMEM = OFF_DW + FD_BYTES;

struct Norm3f { float sub[3], mul[3]; };

struct FrontArgs {
    const void* x;                 // u8 NHWC [B,H,W,3] | fp32 NHWC | fp32 NCHW
    const float* stem_w;           // OIHW [8,3,3,3]
    const float* stem_b;           // [8]
    const __half* same_w;          // VQAE_PACK_SAME_MMA_F16, c = 8: [11][8][8]
    const __half* down_w;          // VQAE_PACK_DOWN_MMA_F16, 8 -> 16
    float* out;                    // NHWC fp32 [B,H/2,W/2,16]
    int n_tiles, H, W, tiles_x, tiles_per_img;
    Norm3f n;
    float s_b1a, s_b1b, s_b2a, s_b2b, s_b3a, s_b3b, s_b4, s_scale;      // 'same' block scalars
    float d_b1a, d_b1b, d_b2a, d_b2b, d_b3a, d_b3b, d_b1c, d_bsum;      // 'down' block scalars
};

// XKIND: 0 = fp32 NCHW, 1 = fp32 NHWC, 2 = u8 NHWC (normalised like normalize_u8 / stem_in)
template <int XKIND>
__device__ __forceinline__ void load_px(const void* xv, int64_t b, int iy, int ix, int H, int W,
                                        const Norm3f& n, float (&v)[3]) {
    const int64_t hw = (int64_t)H * W;
    if (XKIND == 0) {
        const float* s = reinterpret_cast<const float*>(xv) + b * 3 * hw + (int64_t)iy * W + ix;
        v[0] = __ldg(s); v[1] = __ldg(s + hw); v[2] = __ldg(s + 2 * hw);
    } else if (XKIND == 1) {
        const float* s = reinterpret_cast<const float*>(xv) + (b * hw + (int64_t)iy * W + ix) * 3;
        v[0] = __ldg(s); v[1] = __ldg(s + 1); v[2] = __ldg(s + 2);
    } else {
        const uint8_t* s = reinterpret_cast<const uint8_t*>(xv) + (b * hw + (int64_t)iy * W + ix) * 3;
        v[0] = __fmul_rn(__fsub_rn((float)__ldg(s), n.sub[0]), n.mul[0]);
        v[1] = __fmul_rn(__fsub_rn((float)__ldg(s + 1), n.sub[1]), n.mul[1]);
        v[2] = __fmul_rn(__fsub_rn((float)__ldg(s + 2), n.sub[2]), n.mul[2]);
    }
}

__device__ __forceinline__ void split_h2(float f0, float f1, uint32_t& hi, uint32_t& lo) {
    const __half2 hh = __floats2half2_rn(f0, f1);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&hh);
    lo = *reinterpret_cast<const uint32_t*>(&ll);
}

template <int XKIND>
__global__ void __launch_bounds__(FF_THREADS, 3)
front_fused_kernel(FrontArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = tc::smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    // ---- one-time: weights ----
    // stem B fragments (col-major K x 8): k' = kx * 4 + c, zero for kx == 3 or c == 3; hi / lo
    uint32_t sbh[3][2], sbl[3][2];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            // k-slots (2t, 2t + 1) [h = 0] and (2t + 8, 2t + 9) [h = 1] of this lane stand for window
            // pixel kx = t, channels (0, 1) and (2, pad)
            float w[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int kx = t, c = 2 * h + e;
                w[e] = (kx < 3 && c < 3) ? __ldg(a.stem_w + g * 27 + c * 9 + ky * 3 + kx) : 0.f;
            }
            split_h2(w[0], w[1], sbh[ky][h], sbl[ky][h]);
        }
    const float sbias0 = __ldg(a.stem_b + 2 * t), sbias1 = __ldg(a.stem_b + 2 * t + 1);
    // 'same' block: all eleven 8 x 8 matrices as B fragments in registers (mma_same.cu, K8 form)
    uint32_t w1f, w3f, w2f[9];
    {
        auto frag = [&](const __half* w) {
            return __ldg(reinterpret_cast<const uint32_t*>(w + g * 8 + 2 * t));
        };
        w1f = frag(a.same_w);
        w3f = frag(a.same_w + 10 * 64);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) w2f[tap] = frag(a.same_w + (1 + tap) * 64);
    }
    // 'down' block weights -> shared memory rows (mma_down.cu)
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.down_w);
        uint8_t* dw = smem + OFF_DW;
        constexpr int PI = FD_CI / 8, PO = FD_CO / 8;
        constexpr int R1 = FD_CO, R2 = 4 * FD_CO, R3 = FD_CO, RS = 4 * FD_CO;
        for (int i = tid; i < R1 * PI; i += FF_THREADS)
            *reinterpret_cast<uint4*>(dw + FD_OFF_W1 + (i / PI) * FD_WPI + (i % PI) * 16) = __ldg(src + i);
        src += R1 * PI;
        for (int i = tid; i < R2 * PO; i += FF_THREADS)
            *reinterpret_cast<uint4*>(dw + FD_OFF_W2 + (i / PO) * FD_WPO + (i % PO) * 16) = __ldg(src + i);
        src += R2 * PO;
        for (int i = tid; i < R3 * PO; i += FF_THREADS)
            *reinterpret_cast<uint4*>(dw + FD_OFF_W3 + (i / PO) * FD_WPO + (i % PO) * 16) = __ldg(src + i);
        src += R3 * PO;
        for (int i = tid; i < RS * PI; i += FF_THREADS)
            *reinterpret_cast<uint4*>(dw + FD_OFF_WS + (i / PI) * FD_WPI + (i % PI) * 16) = __ldg(src + i);
    }
    // staged planes: zero once (the two padding pixels per row and the slack row are never written)
    for (int i = tid; i < (int)(S_BYTES / 16); i += FF_THREADS)
        *reinterpret_cast<uint4*>(smem + OFF_S + i * 16) = make_uint4(0, 0, 0, 0);
    // slack rows of x0 / U (q >= 612) stay finite
    for (int i = tid; i < (int)((OFF_DW - OFF_XS) / 16); i += FF_THREADS)
        *reinterpret_cast<uint4*>(smem + OFF_XS + i * 16) = make_uint4(0, 0, 0, 0);

    const ActC sact1(a.s_b1a, a.s_b1b), sact2(a.s_b2a, a.s_b2b), sact3(a.s_b3a, a.s_b3b);
    const ActC dact1(a.d_b1a, a.d_b1b), dact2(a.d_b2a, a.d_b2b), dact3(a.d_b3a, a.d_b3b);
    const uint32_t sU = sbase + OFF_U, sDW = sbase + OFF_DW;
    const int Wo = a.W / 2;
    __syncthreads();

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int img = tile / a.tiles_per_img;
        const int trem = tile - img * a.tiles_per_img;
        const int r0 = (trem / a.tiles_x) * FF_TH, c0 = (trem % a.tiles_x) * FF_TW;
        const bool top = r0 == 0, bottom = r0 + FF_TH == a.H, left = c0 == 0, right = c0 + FF_TW == a.W;
        const bool border = top || bottom || left || right;

        // ================= stage the input window: 20 x 36 pixels, zero outside the image ==========
        for (int i = tid; i < FF_SREAL_R * FF_SREAL_C; i += FF_THREADS) {
            const int si = i / FF_SREAL_C, sj = i - si * FF_SREAL_C;
            const int iy = r0 - 2 + si, ix = c0 - 2 + sj;
            float v[3] = {0.f, 0.f, 0.f};
            if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) load_px<XKIND>(a.x, img, iy, ix, a.H, a.W, a.n, v);
            uint4 px;                                      // hi01, hi2-, lo01, lo2-
            split_h2(v[0], v[1], px.x, px.z);
            split_h2(v[2], 0.f, px.y, px.w);
            *reinterpret_cast<uint4*>(smem + OFF_S + (uint32_t)(si * FF_SPW + sj) * 16) = px;
        }
        __syncthreads();

        // ================= stem: x0[q] for the 18 x 34 ring'd tile =================================
        for (int m = warp; m < FF_MT1; m += FF_WARPS) {
            const int q0 = 16 * m + g, q1 = q0 + 8;
            const int lr0 = q0 / FF_PW, lc0 = q0 - lr0 * FF_PW;
            const int lr1 = q1 / FF_PW, lc1 = q1 - lr1 * FF_PW;
            // pixel (lr, lc) = staged (lr + 1, lc + 1); its window starts at staged (lr + ky, lc);
            // this lane's window pixel is kx = t
            const uint8_t* p0 = smem + OFF_S + (uint32_t)(lr0 * FF_SPW + lc0 + t) * 16;
            const uint8_t* p1 = smem + OFF_S + (uint32_t)(lr1 * FF_SPW + lc1 + t) * 16;
            float d[4] = {sbias0, sbias1, sbias0, sbias1};
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const uint4 u0 = *reinterpret_cast<const uint4*>(p0 + ky * FF_SPW * 16);
                const uint4 u1 = *reinterpret_cast<const uint4*>(p1 + ky * FF_SPW * 16);
                const uint32_t ah[4] = {u0.x, u1.x, u0.y, u1.y}, al[4] = {u0.z, u1.z, u0.w, u1.w};
                mma_16816(d, al, sbh[ky][0], sbh[ky][1]);
                mma_16816(d, ah, sbl[ky][0], sbl[ky][1]);
                mma_16816(d, ah, sbh[ky][0], sbh[ky][1]);
            }
            // ring pixels that wrap around the image are written by the border pass below
            const bool w0 = border && ((top && lr0 == 0) || (bottom && lr0 == FF_TH + 1) || (left && lc0 == 0) ||
                            (right && lc0 == FF_PW - 1));
            const bool w1 = border && ((top && lr1 == 0) || (bottom && lr1 == FF_TH + 1) || (left && lc1 == 0) ||
                            (right && lc1 == FF_PW - 1));
            if (!w0) *reinterpret_cast<float2*>(smem + OFF_XS + q0 * 32 + 8 * t) = make_float2(d[0], d[1]);
            if (!w1) *reinterpret_cast<float2*>(smem + OFF_XS + q1 * 32 + 8 * t) = make_float2(d[2], d[3]);
        }
        // border pass: x0 of the wrapped ring pixels at their image coordinates (zero padding there),
        // one (pixel, output channel) per thread and step; ring index: top row, bottom row, left, right
        if (border) {
            constexpr int NRING = 2 * FF_PW + 2 * FF_TH;
            for (int i = tid; i < NRING * 8; i += FF_THREADS) {
                const int ri = i >> 3, o = i & 7;
                int lr, lc;
                if (ri < FF_PW) { lr = 0; lc = ri; }
                else if (ri < 2 * FF_PW) { lr = FF_TH + 1; lc = ri - FF_PW; }
                else if (ri < 2 * FF_PW + FF_TH) { lr = ri - 2 * FF_PW + 1; lc = 0; }
                else { lr = ri - 2 * FF_PW - FF_TH + 1; lc = FF_PW - 1; }
                const bool wr = (top && lr == 0) || (bottom && lr == FF_TH + 1) || (left && lc == 0) ||
                                (right && lc == FF_PW - 1);
                if (!wr) continue;
                int y = r0 - 1 + lr, x = c0 - 1 + lc;
                y = y < 0 ? y + a.H : (y >= a.H ? y - a.H : y);
                x = x < 0 ? x + a.W : (x >= a.W ? x - a.W : x);
                float acc = __ldg(a.stem_b + o);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int iy = y + ky - 1;
                    if (iy < 0 || iy >= a.H) continue;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int ix = x + kx - 1;
                        if (ix < 0 || ix >= a.W) continue;
                        float v[3];
                        load_px<XKIND>(a.x, img, iy, ix, a.H, a.W, a.n, v);
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            acc = fmaf(v[c], __ldg(a.stem_w + o * 27 + c * 9 + ky * 3 + kx), acc);
                    }
                }
                reinterpret_cast<float*>(smem + OFF_XS)[(lr * FF_PW + lc) * 8 + o] = acc;
            }
        }
        __syncthreads();

        // ================= 'same' block, stage 1: U = f16(elu(W1 . f16(elu(x0 + b1a) + b1b) + b2a) + b2b)
        const uint8_t* xs = smem + OFF_XS;
        for (int m = warp; m < FF_MT1; m += FF_WARPS) {
            const int q0 = 16 * m + g, q1 = q0 + 8;
            const float2 v0 = *reinterpret_cast<const float2*>(xs + q0 * 32 + 8 * t);
            const float2 v1 = *reinterpret_cast<const float2*>(xs + q1 * 32 + 8 * t);
            const uint32_t a0 = sact1(v0.x, v0.y), a1 = sact1(v1.x, v1.y);
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            mma_1688(d, a0, a1, w1f);
            *reinterpret_cast<uint32_t*>(smem + OFF_U + q0 * 16 + 4 * t) = sact2(d[0], d[1]);
            *reinterpret_cast<uint32_t*>(smem + OFF_U + q1 * 16 + 4 * t) = sact2(d[2], d[3]);
        }
        __syncthreads();

        // ================= 'same' block, stages 2 + 3: y = x0 + scale * W3 . V + b4 -> shared memory
        float* ys = reinterpret_cast<float*>(smem + OFF_S);
        for (int mt = warp; mt < FF_MT2; mt += FF_WARPS) {
            const int r = mt >> 1, cb = (mt & 1) * 16;
            const int qc = (r + 1) * FF_PW + cb + 1;
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;
            const uint32_t lbase = sU + (uint32_t)(qc + lrow) * 16;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int shift = (tap / 3 - 1) * FF_PW + (tap % 3 - 1);
                uint32_t a0, a1;
                ldmatrix_x2(a0, a1, lbase + shift * 16);
                mma_1688(d, a0, a1, w2f[tap]);
            }
            const uint32_t v0 = sact3(d[0], d[1]), v1 = sact3(d[2], d[3]);
            d[0] = d[1] = d[2] = d[3] = 0.f;
            mma_1688(d, v0, v1, w3f);
            const int qa = qc + g, qb = qa + 8;
            const float2 x0 = *reinterpret_cast<const float2*>(xs + qa * 32 + 8 * t);
            const float2 x1 = *reinterpret_cast<const float2*>(xs + qb * 32 + 8 * t);
            float2 o0, o1;
            o0.x = fmaf(d[0], a.s_scale, a.s_b4) + x0.x;
            o0.y = fmaf(d[1], a.s_scale, a.s_b4) + x0.y;
            o1.x = fmaf(d[2], a.s_scale, a.s_b4) + x1.x;
            o1.y = fmaf(d[3], a.s_scale, a.s_b4) + x1.y;
            const int p = r * FF_TW + cb + g;
            *reinterpret_cast<float2*>(ys + p * 8 + 2 * t) = o0;
            *reinterpret_cast<float2*>(ys + (p + 8) * 8 + 2 * t) = o1;
        }
        __syncthreads();

        // ================= 'down' block 8 -> 16 on the tile: 8 output rows x 16 columns, one M-tile
        //                   (output row) per warp; arithmetic of mma_down.cu, CI = 8 ===================
        {
            constexpr int NT = FD_CO / 8, KSO = FD_CO / 16;
            const uint32_t lo_i = (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8) * FD_WPI;
            const uint32_t lo_o = (uint32_t)((lane & 7) + (lane >> 4) * 8) * FD_WPO + ((lane >> 3) & 1) * 16;
            const int orow = warp;
            float d2[NT][4], d3[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) d2[j][e] = d3[j][e] = 0.f;
            const float2 c1c = make_float2(a.d_b1c, a.d_b1c);
#pragma unroll
            for (int pos = 0; pos < 4; ++pos) {
                const int py = 2 * orow + (pos >> 1), px = 2 * g + (pos & 1);
                const float2 u0 = *reinterpret_cast<const float2*>(ys + (py * FF_TW + px) * 8 + 2 * t);
                const float2 u1 = *reinterpret_cast<const float2*>(ys + (py * FF_TW + px + 16) * 8 + 2 * t);
                const uint32_t a10 = dact1(u0.x, u0.y), a11 = dact1(u1.x, u1.y);
                const float2 s0 = __fadd2_rn(u0, c1c), s1 = __fadd2_rn(u1, c1c);
                const uint32_t as0 = pack_h2(s0.x, s0.y), as1 = pack_h2(s1.x, s1.y);
                float d1[NT][4];
#pragma unroll
                for (int j = 0; j < NT; ++j) d1[j][0] = d1[j][1] = d1[j][2] = d1[j][3] = 0.f;
                {
                    uint32_t b0, b1;
                    ldmatrix_x2(b0, b1, sDW + FD_OFF_W1 + lo_i);
                    mma_1688(d1[0], a10, a11, b0);
                    mma_1688(d1[1], a10, a11, b1);
                }
                uint32_t uf[KSO][4];
                uf[0][0] = dact2(d1[0][0], d1[0][1]);
                uf[0][1] = dact2(d1[0][2], d1[0][3]);
                uf[0][2] = dact2(d1[1][0], d1[1][1]);
                uf[0][3] = dact2(d1[1][2], d1[1][3]);
                {
                    uint32_t bf[4];
                    ldmatrix_x4(bf, sDW + FD_OFF_W2 + (uint32_t)(pos * FD_CO) * FD_WPO + lo_o);
                    mma_16816(d2[0], uf[0], bf[0], bf[1]);
                    mma_16816(d2[1], uf[0], bf[2], bf[3]);
                }
                {
                    uint32_t b0, b1;
                    ldmatrix_x2(b0, b1, sDW + FD_OFF_WS + (uint32_t)(pos * FD_CO) * FD_WPI + lo_i);
                    mma_1688(d3[0], as0, as1, b0);
                    mma_1688(d3[1], as0, as1, b1);
                }
            }
            {
                uint32_t vf[KSO][4];
                vf[0][0] = dact3(d2[0][0], d2[0][1]);
                vf[0][1] = dact3(d2[0][2], d2[0][3]);
                vf[0][2] = dact3(d2[1][0], d2[1][1]);
                vf[0][3] = dact3(d2[1][2], d2[1][3]);
                uint32_t bf[4];
                ldmatrix_x4(bf, sDW + FD_OFF_W3 + lo_o);
                mma_16816(d3[0], vf[0], bf[0], bf[1]);
                mma_16816(d3[1], vf[0], bf[2], bf[3]);
            }
            float* o0 = a.out + (size_t)img * (a.H / 2) * Wo * FD_CO +
                        ((size_t)(r0 / 2 + orow) * Wo + c0 / 2 + g) * FD_CO + 4 * t;
            float* o1 = o0 + 8 * FD_CO;
            const float (&e)[4] = d3[0], (&f)[4] = d3[1];
            *reinterpret_cast<float4*>(o0) =
                make_float4(e[0] + a.d_bsum, e[1] + a.d_bsum, f[0] + a.d_bsum, f[1] + a.d_bsum);
            *reinterpret_cast<float4*>(o1) =
                make_float4(e[2] + a.d_bsum, e[3] + a.d_bsum, f[2] + a.d_bsum, f[3] + a.d_bsum);
        }
        __syncthreads();                // y (aliases the staged planes), x0 and U are free again
        // the staged planes' padding pixels and slack row were overwritten by y: zero them again
        for (int i = tid; i < FF_SROWS * 2 + FF_SPW; i += FF_THREADS) {
            // two padding pixels per row, and the whole slack row
            uint32_t off;
            if (i < FF_SROWS * 2) off = (uint32_t)((i >> 1) * FF_SPW + FF_SREAL_C + (i & 1)) * 16;
            else off = (uint32_t)((FF_SROWS - 1) * FF_SPW + (i - FF_SROWS * 2)) * 16;
            *reinterpret_cast<uint4*>(smem + OFF_S + off) = make_uint4(0, 0, 0, 0);
        }
    }
}

template <int XKIND>
int launch_front(const FrontArgs& a, int sm_count, cudaStream_t stream) {
    auto kern = front_fused_kernel<XKIND>;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM));
        attr_set.cur() = true;
    }
    const int cap = sm_count * 3;
    const int grid = a.n_tiles < cap ? a.n_tiles : cap;
    kern<<<grid, FF_THREADS, FF_SMEM, stream>>>(a);
    return check_launch();
}

}  // namespace

bool front_fused_supported(int H, int W) {
    return H >= FF_TH && W >= FF_TW && H % FF_TH == 0 && W % FF_TW == 0;
}

int front_fused(const void* x, int x_dtype, int x_layout, const float* stem_w, const float* stem_b,
                const float* mean, const float* stdv, const void* same_w, const float* same_scalars8,
                const void* down_w, const float* down_scalars8, float* out, int64_t B, int H, int W,
                int sm_count, cudaStream_t stream) {
    if (!x || !stem_w || !stem_b || !same_w || !same_scalars8 || !down_w || !down_scalars8 || !out ||
        B <= 0)
        return VQAE_ERR_BAD_ARG;
    if (!front_fused_supported(H, W)) return VQAE_ERR_UNSUPPORTED;
    FrontArgs a{};
    a.x = x; a.stem_w = stem_w; a.stem_b = stem_b;
    a.same_w = reinterpret_cast<const __half*>(same_w);
    a.down_w = reinterpret_cast<const __half*>(down_w);
    a.out = out; a.H = H; a.W = W;
    a.tiles_x = W / FF_TW; a.tiles_per_img = (H / FF_TH) * a.tiles_x;
    const int64_t nt = B * a.tiles_per_img;
    if (nt > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    a.n_tiles = (int)nt;
    const float* s = same_scalars8;
    a.s_b1a = s[0]; a.s_b1b = s[1]; a.s_b2a = s[2]; a.s_b2b = s[3]; a.s_b3a = s[4]; a.s_b3b = s[5];
    a.s_b4 = s[6]; a.s_scale = s[7];
    const float* d = down_scalars8;
    a.d_b1a = d[0]; a.d_b1b = d[1]; a.d_b2a = d[2]; a.d_b2b = d[3]; a.d_b3a = d[4]; a.d_b3b = d[5];
    a.d_b1c = d[6]; a.d_bsum = d[7];
    int kind;
    if (x_dtype == VQAE_DT_U8) {
        if (!mean || !stdv) return VQAE_ERR_BAD_ARG;
        if (x_layout != VQAE_LAYOUT_NHWC) return VQAE_ERR_UNSUPPORTED;
        for (int c = 0; c < 3; ++c) {
            a.n.sub[c] = mean[c] * 255.0f;
            a.n.mul[c] = 1.0f / (stdv[c] * 255.0f);
        }
        kind = 2;
    } else if (x_dtype == VQAE_DT_F32) {
        kind = x_layout == VQAE_LAYOUT_NCHW ? 0 : 1;
    } else {
        return VQAE_ERR_UNSUPPORTED;
    }
    if (kind == 2) return launch_front<2>(a, sm_count, stream);
    if (kind == 0) return launch_front<0>(a, sm_count, stream);
    return launch_front<1>(a, sm_count, stream);
}

}  // namespace vqae
