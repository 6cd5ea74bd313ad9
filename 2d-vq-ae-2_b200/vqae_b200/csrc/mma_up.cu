// PreActFixupResBlock 'up' (c_in -> c_in / 2, x2 bicubic; layers/conv_block.py:196-216 with
// ResizeConv2D, layers/conv.py:4-11) on warp-level tensor-core MMAs, in two kernels.  A 1x1 conv
// without bias commutes with the (linear) upsample, so both ResizeConv2Ds run their conv at LOW
// resolution and only the interpolation + branch_conv3 run at high resolution:
//
//   head (low resolution, pointwise, register-resident like mma_down.cu)
//       A1 = f16(elu(x + b1a) + b1b);  D1 = A1 . W1^T;  U = f16(elu(D1 + b2a) + b2b)
//       T2 = U . W2^T                       -> fp32 [B,H,W,CB]      (branch, before the upsample)
//       S  = f16(x + b1c) . Ws^T            -> fp32 [B,H,W,CO]      (skip, before the upsample)
//   tail (high resolution, CTA = 8 or 16 x 32 output pixels)
//       low-res window of T2 | S (index-clamped like nn.Upsample(bicubic, align_corners=False)) staged
//       in shared memory as split fp16 planes; per warp and 16-pixel row segment the horizontal 4-tap
//       pass is a tensor-core GEMM with the constant tap matrix (cubic taps for scale 2:
//       (-9, 67, 225, -27) / 256 and its mirror), the vertical pass an fp32 combination of the four
//       source rows' accumulators -- the result IS the A fragment (branch) / output summand (skip)
//       V = f16(elu(t3 + b3a) + b3b);  D3 = V . (scale W3)^T
//       out = D3 + skip + (b4 + b1d)       fp32 NHWC, 128-bit stores
//
// Channel orders of the weights are permuted at pack time (pack.cu, VQAE_PACK_UP_MMA_F16) so that a
// lane's fragment slots are four consecutive channels in memory, as in mma_same.cu.
#include "common.cuh"
#include "kernels.cuh"
#include "mma_common.cuh"

namespace vqae {
namespace {

using namespace mma;

constexpr int MU_WARPS = 8;
constexpr int MU_THREADS = MU_WARPS * 32;

// ------------------------------------------------------------------------------------------------
// head
// ------------------------------------------------------------------------------------------------
template <int CI>
struct UhCfg {
    static constexpr int CB = CI, CO = CI / 2;
    static constexpr int KS = CI / 16, NTB = CB / 8, NTO = CO / 8;
    static constexpr int WP = CI * 2 + 16;                 // row pitch of every matrix (K = CI)
    static constexpr uint32_t OFF_W1 = 0;
    static constexpr uint32_t OFF_W2 = OFF_W1 + CB * WP;
    static constexpr uint32_t OFF_WS = OFF_W2 + CB * WP;
    static constexpr uint32_t SMEM = OFF_WS + CO * WP;
    static constexpr int MIN_CTAS = CI <= 32 ? 2 : 1;
};

struct UhArgs {
    const float* x;       // [P][CI] fp32 (NHWC, pixels flattened)
    float* t2;            // [P][CB]
    float* s;             // [P][CO]
    const __half* w;      // [W1 | W2 | Ws | W3] dense [n][k] fp16 (pack.cu)
    int n_mtiles;         // ceil(P / 16)
    int64_t P;
    float b1a, b1b, b2a, b2b, b1c;
};

template <int CI>
__global__ void __launch_bounds__(MU_THREADS, UhCfg<CI>::MIN_CTAS)
up_head_mma_kernel(UhArgs a) {
    using Cfg = UhCfg<CI>;
    constexpr int CB = Cfg::CB, CO = Cfg::CO, KS = Cfg::KS, NTB = Cfg::NTB, NTO = Cfg::NTO;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = tc::smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.w);
        constexpr int PP = CI / 8;                                     // 16-byte pieces per row
        for (int i = tid; i < (2 * CB + CO) * PP; i += MU_THREADS)
            *reinterpret_cast<uint4*>(smem + (i / PP) * Cfg::WP + (i % PP) * 16) = __ldg(src + i);
    }
    __syncthreads();
    const ActC act1(a.b1a, a.b1b), act2(a.b2a, a.b2b);
    const float2 c1c = make_float2(a.b1c, a.b1c);
    const uint32_t lo = (uint32_t)((lane & 7) + (lane >> 4) * 8) * Cfg::WP + ((lane >> 3) & 1) * 16;

    for (int m = blockIdx.x * MU_WARPS + warp; m < a.n_mtiles; m += gridDim.x * MU_WARPS) {
        const int64_t p0 = (int64_t)m * 16 + g, p1 = p0 + 8;
        const int64_t q0 = p0 < a.P ? p0 : a.P - 1, q1 = p1 < a.P ? p1 : a.P - 1;
        uint32_t a1[KS][4], as[KS][4];
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(a.x + q0 * CI + 16 * s + 4 * t));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(a.x + q1 * CI + 16 * s + 4 * t));
            a1[s][0] = act1(v0.x, v0.y);
            a1[s][1] = act1(v1.x, v1.y);
            a1[s][2] = act1(v0.z, v0.w);
            a1[s][3] = act1(v1.z, v1.w);
            const float2 s0 = __fadd2_rn(make_float2(v0.x, v0.y), c1c), s1 = __fadd2_rn(make_float2(v1.x, v1.y), c1c);
            const float2 s2 = __fadd2_rn(make_float2(v0.z, v0.w), c1c), s3 = __fadd2_rn(make_float2(v1.z, v1.w), c1c);
            as[s][0] = pack_h2(s0.x, s0.y);
            as[s][1] = pack_h2(s1.x, s1.y);
            as[s][2] = pack_h2(s2.x, s2.y);
            as[s][3] = pack_h2(s3.x, s3.y);
        }
        // ---- skip: S = As . Ws^T ----
        {
            float d[NTO][4];
#pragma unroll
            for (int j = 0; j < NTO; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
            if constexpr (NTO == 1) {
                // c_out = 8: one n-tile, B fragments by two 32-bit reads per k-step
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    const uint8_t* row = smem + Cfg::OFF_WS + g * Cfg::WP + (16 * s + 2 * t) * 2;
                    mma_16816(d[0], as[s], *reinterpret_cast<const uint32_t*>(row),
                              *reinterpret_cast<const uint32_t*>(row + 16));
                }
                if (p0 < a.P) *reinterpret_cast<float2*>(a.s + p0 * CO + 2 * t) = make_float2(d[0][0], d[0][1]);
                if (p1 < a.P) *reinterpret_cast<float2*>(a.s + p1 * CO + 2 * t) = make_float2(d[0][2], d[0][3]);
            } else {
#pragma unroll
                for (int p = 0; p < NTO / 2; ++p)
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        uint32_t bf[4];
                        ldmatrix_x4(bf, sbase + Cfg::OFF_WS + lo + (uint32_t)(16 * p) * Cfg::WP + s * 32);
                        mma_16816(d[2 * p], as[s], bf[0], bf[1]);
                        mma_16816(d[2 * p + 1], as[s], bf[2], bf[3]);
                    }
#pragma unroll
                for (int p = 0; p < NTO / 2; ++p) {
                    const float (&e)[4] = d[2 * p], (&f)[4] = d[2 * p + 1];
                    if (p0 < a.P) *reinterpret_cast<float4*>(a.s + p0 * CO + 16 * p + 4 * t) = make_float4(e[0], e[1], f[0], f[1]);
                    if (p1 < a.P) *reinterpret_cast<float4*>(a.s + p1 * CO + 16 * p + 4 * t) = make_float4(e[2], e[3], f[2], f[3]);
                }
            }
        }
        // ---- branch: D1 = A1 . W1^T, U = act2(D1), T2 = U . W2^T ----
        uint32_t uf[KS][4];
        {
            float d[NTB][4];
#pragma unroll
            for (int j = 0; j < NTB; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
#pragma unroll
            for (int p = 0; p < NTB / 2; ++p)
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    uint32_t bf[4];
                    ldmatrix_x4(bf, sbase + Cfg::OFF_W1 + lo + (uint32_t)(16 * p) * Cfg::WP + s * 32);
                    mma_16816(d[2 * p], a1[s], bf[0], bf[1]);
                    mma_16816(d[2 * p + 1], a1[s], bf[2], bf[3]);
                }
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                uf[s][0] = act2(d[2 * s][0], d[2 * s][1]);
                uf[s][1] = act2(d[2 * s][2], d[2 * s][3]);
                uf[s][2] = act2(d[2 * s + 1][0], d[2 * s + 1][1]);
                uf[s][3] = act2(d[2 * s + 1][2], d[2 * s + 1][3]);
            }
        }
        // T2 in n-tile pairs (16 output channels at a time keeps the accumulators small)
#pragma unroll
        for (int p = 0; p < NTB / 2; ++p) {
            float e[4] = {0.f, 0.f, 0.f, 0.f}, f[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                uint32_t bf[4];
                ldmatrix_x4(bf, sbase + Cfg::OFF_W2 + lo + (uint32_t)(16 * p) * Cfg::WP + s * 32);
                mma_16816(e, uf[s], bf[0], bf[1]);
                mma_16816(f, uf[s], bf[2], bf[3]);
            }
            if (p0 < a.P) *reinterpret_cast<float4*>(a.t2 + p0 * CB + 16 * p + 4 * t) = make_float4(e[0], e[1], f[0], f[1]);
            if (p1 < a.P) *reinterpret_cast<float4*>(a.t2 + p1 * CB + 16 * p + 4 * t) = make_float4(e[2], e[3], f[2], f[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tail
// ------------------------------------------------------------------------------------------------
struct UtArgs {
    const float* t2;      // [B,H,W,CB]
    const float* s;       // [B,H,W,CO]
    const __half* w3;     // [CO][CB] fp16, scale folded in, channel orders permuted (pack.cu)
    float* out;           // [B,2H,2W,CO]
    int H, W;             // low-res extent
    int tiles_x, tiles_per_img, n_tiles;
    FastDiv fd_img, fd_tx;
    float b3a, b3b, bsum; // bsum = b4 + b1d
};

// ------------------------------------------------------------------------------------------------
// interpolation on the tensor cores (round 2)
// ------------------------------------------------------------------------------------------------
// The first form of this kernel (separable bicubic with fp32 FMAs: a vertical 4-tap pass into a
// warp-private buffer, then a horizontal pass into the fragments) was bound by its shared-memory reads
// (ncu: L1 data pipe 88 %: ~28 128-bit loads per lane and 16 output pixels; 1.0 ms for the C = 16 block
// at batch 256 against 0.70 ms now).  The x2 interpolation is a constant linear
// map, so here the HORIZONTAL pass is a GEMM:  for every source row r of the 4-row footprint
//      D_r[16 px][ch] = Hx[16 px][12 (+4) source cols] . X_r[source cols][ch]
// with Hx the tap matrix (entries k / 256: exact in fp16) as the A operand, and the staged window --
// split into fp16 hi + lo planes when it is staged, stored [row][col][slot] and read with transposed
// ldmatrix (a matrix row = eight consecutive slots of one source pixel) -- as the B operand; the vertical pass
// is  sum_r wy[r] D_r  on the accumulators (fp32).  The result is in accumulator-fragment layout, i.e.
// already the A fragment of branch_conv3 (branch channels) or the summand of the output fragment
// (skip channels): no intermediate buffer.  12 ldmatrix + 24 small MMAs per 16 pixels at C = 16.
// Channels are staged in fragment-slot order (slot s holds channel perm(s), pack.cu), matching W3.
constexpr int U3_TX = 32, U3_RX = U3_TX / 2 + 4;                   // 20 window columns

// TY: output rows per CTA tile (8 or 16; the taller tile stages 12 window rows for 512 pixels instead
// of 8 for 256 -- a quarter fewer staged pixels per output pixel and half as many CTA barriers)
template <int CB, int TY>
struct U3Cfg {
    static constexpr int U3_TY = TY, U3_RY = TY / 2 + 4;
    static constexpr int CO = CB / 2, CT = CB + CO;
    static constexpr int KS = CB / 16, NTO = CO / 8;
    static constexpr int WP = CB * 2 + 16;                        // W3 row pitch
    // staged window: [row][col][slot] fp16, a pixel = CT halves padded to an odd number of 16-byte
    // groups (the eight pixel rows of an ldmatrix then fall into distinct bank groups)
    static constexpr int PIXB = CT * 2 + (((CT / 8) & 1) ? 0 : 16);
    static constexpr uint32_t PLANE = (U3_RY * U3_RX + 4) * PIXB;             // + 4 pixels of over-read
    static constexpr uint32_t OFF_HI = 0, OFF_LO = PLANE, OFF_W3 = 2 * PLANE;
    static constexpr uint32_t SMEM = OFF_W3 + CO * WP;
    static constexpr int NITEMS = U3_RY * U3_RX * (CT / 4);                  // float4 pieces of a window
    static constexpr int NI = (NITEMS + MU_THREADS - 1) / MU_THREADS;
    static constexpr bool PREFETCH = NI <= 6;                     // next tile's pieces held in registers
    static constexpr int MIN_CTAS = CB <= 32 ? 3 : (CB == 64 ? 2 : 1);
};

template <int CB, int TY>
__global__ void __launch_bounds__(MU_THREADS, U3Cfg<CB, TY>::MIN_CTAS)
up_tail3_mma_kernel(UtArgs a) {
    using Cfg = U3Cfg<CB, TY>;
    constexpr int U3_TY = Cfg::U3_TY, U3_RY = Cfg::U3_RY;
    constexpr int CO = Cfg::CO, CT = Cfg::CT, KS = Cfg::KS, NTO = Cfg::NTO, PIXB = Cfg::PIXB, NI = Cfg::NI;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = tc::smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.w3);
        constexpr int PP = CB / 8;
        for (int i = tid; i < CO * PP; i += MU_THREADS)
            *reinterpret_cast<uint4*>(smem + Cfg::OFF_W3 + (i / PP) * Cfg::WP + (i % PP) * 16) = __ldg(src + i);
        for (int i = tid; i < (int)(2 * Cfg::PLANE / 16); i += MU_THREADS)       // padding stays zero
            *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
    }
    const ActC act3(a.b3a, a.b3b);
    const uint32_t lo3 = (uint32_t)((lane & 7) + (lane >> 4) * 8) * Cfg::WP + ((lane >> 3) & 1) * 16;
    const int Ho = 2 * a.H, Wo = 2 * a.W;
    constexpr float WE[4] = {-9.f / 256.f, 67.f / 256.f, 225.f / 256.f, -27.f / 256.f};    // even index
    constexpr float WO[4] = {-27.f / 256.f, 225.f / 256.f, 67.f / 256.f, -9.f / 256.f};    // odd index
    // A fragment of the horizontal tap matrix Hx[i][k]: output pixel i = g (+ 8) of the segment reads
    // source columns vc0(i) .. vc0(i) + 3, vc0(i) = (i >> 1) + (i & 1), taps by the parity of i
    uint32_t hx[4];
    {
        auto tap = [&](int i, int k) {
            const int d = k - ((i >> 1) + (i & 1));
            float w = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) w = d == e ? ((i & 1) ? WO[e] : WE[e]) : w;
            return w;
        };
        hx[0] = pack_h2(tap(g, 2 * t), tap(g, 2 * t + 1));
        hx[1] = pack_h2(tap(g + 8, 2 * t), tap(g + 8, 2 * t + 1));
        hx[2] = pack_h2(tap(g, 2 * t + 8), tap(g, 2 * t + 9));
        hx[3] = pack_h2(tap(g + 8, 2 * t + 8), tap(g + 8, 2 * t + 9));
    }
    // transposed ldmatrix: matrix m = lane >> 3 = (n-tile of the pair, k half); its row lane & 7 is the
    // source column (pixel) 8 * (k half) + (lane & 7), eight consecutive slots of that pixel
    const uint32_t lm = (uint32_t)((((lane >> 3) & 1) * 8 + (lane & 7)) * PIXB + (lane >> 4) * 16);

    // window pieces of this thread: item i = tid + k * 256 -> (channel quad, window pixel); the pieces of
    // the NEXT tile are loaded into registers before the current tile is computed
    float4 pv[NI];
    auto fetch = [&](int tile) {
        const int img = a.fd_img.d == 1 ? tile : a.fd_img.div(tile);
        const int rem = tile - img * a.tiles_per_img;
        const int ty = a.fd_tx.d == 1 ? rem : a.fd_tx.div(rem);
        const int ly0 = ty * (U3_TY / 2) - 2, lx0 = (rem - ty * a.tiles_x) * (U3_TX / 2) - 2;
        const float* t2 = a.t2 + (size_t)img * a.H * a.W * CB;
        const float* sk = a.s + (size_t)img * a.H * a.W * CO;
#pragma unroll
        for (int k = 0; k < NI; ++k) {
            const int i = tid + k * MU_THREADS;
            if (i < Cfg::NITEMS) {
                const int quad = i / (U3_RY * U3_RX), p = i - quad * (U3_RY * U3_RX);
                const int ry = p / U3_RX, rx = p - ry * U3_RX;
                const int y = min(max(ly0 + ry, 0), a.H - 1), x = min(max(lx0 + rx, 0), a.W - 1);
                pv[k] = __ldg(reinterpret_cast<const float4*>(
                    quad < CB / 4 ? t2 + ((size_t)y * a.W + x) * CB + 4 * quad
                                  : sk + ((size_t)y * a.W + x) * CO + 4 * (quad - CB / 4)));
            }
        }
    };
    if (Cfg::PREFETCH && (int)blockIdx.x < a.n_tiles) fetch(blockIdx.x);

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int img = a.fd_img.d == 1 ? tile : a.fd_img.div(tile);
        const int rem = tile - img * a.tiles_per_img;
        const int ty = a.fd_tx.d == 1 ? rem : a.fd_tx.div(rem);
        const int oy0 = ty * U3_TY, ox0 = (rem - ty * a.tiles_x) * U3_TX;
        const int ly0 = oy0 / 2 - 2;

        if (!Cfg::PREFETCH) fetch(tile);
        __syncthreads();                                          // previous tile's window reads done
        // ---- stage the T2 | S window as fp16 hi / lo planes [row][col][slot]: the four channels of a
        //      piece go to slots s, s + 1, s + 8, s + 9 of their 16-group (fragment-slot order, pack.cu)
#pragma unroll
        for (int k = 0; k < NI; ++k) {
            const int i = tid + k * MU_THREADS;
            if (i < Cfg::NITEMS) {
                const int quad = i / (U3_RY * U3_RX), p = i - quad * (U3_RY * U3_RX);
                const bool br = quad < CB / 4;
                const int c0 = br ? 4 * quad : 4 * (quad - CB / 4);
                const bool ident = !br && CO < 16;                // an 8-channel group is not permuted
                const int s0 = (br ? 0 : CB) + (ident ? c0 : (c0 & ~15) + 2 * ((c0 >> 2) & 3));
                const float4 v = pv[k];
                const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y);
                const __half2 l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
                uint8_t* d = smem + p * PIXB + s0 * 2;
                const int o2 = ident ? 4 : 16;                    // byte offset of the second slot pair
                *reinterpret_cast<__half2*>(d + Cfg::OFF_HI) = h01;
                *reinterpret_cast<__half2*>(d + Cfg::OFF_HI + o2) = h23;
                *reinterpret_cast<__half2*>(d + Cfg::OFF_LO) = l01;
                *reinterpret_cast<__half2*>(d + Cfg::OFF_LO + o2) = l23;
            }
        }
        __syncthreads();
        if (Cfg::PREFETCH && tile + (int)gridDim.x < a.n_tiles) fetch(tile + gridDim.x);

        // ---- 16 M-tiles (8 rows x 2 segments of 16 pixels), two per warp ----
#pragma unroll 1
        for (int mt = warp; mt < 2 * U3_TY; mt += MU_WARPS) {
            const int ry_o = mt >> 1, seg = mt & 1;
            const int oy = oy0 + ry_o;
            const int wy0 = (oy >> 1) - 2 + (oy & 1) - ly0;       // window-local first source row
            float wy[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) wy[k] = (oy & 1) ? WO[k] : WE[k];
            const uint32_t wbase = sbase + (uint32_t)((wy0 * U3_RX + seg * 8) * PIXB) + lm;
            // interpolated values of the n-tile pair starting at staged slot s0 (16 channels)
            auto interp_pair = [&](int s0, float (&o)[2][4]) {
                float dr[4][2][4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) dr[r][j][0] = dr[r][j][1] = dr[r][j][2] = dr[r][j][3] = 0.f;
                    uint32_t bh[4], bl[4];
                    ldmatrix_x4_trans(bh, wbase + Cfg::OFF_HI + (uint32_t)(r * U3_RX * PIXB + s0 * 2));
                    ldmatrix_x4_trans(bl, wbase + Cfg::OFF_LO + (uint32_t)(r * U3_RX * PIXB + s0 * 2));
                    mma_16816(dr[r][0], hx, bl[0], bl[1]);
                    mma_16816(dr[r][1], hx, bl[2], bl[3]);
                    mma_16816(dr[r][0], hx, bh[0], bh[1]);
                    mma_16816(dr[r][1], hx, bh[2], bh[3]);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        o[j][e] = fmaf(dr[3][j][e], wy[3], fmaf(dr[2][j][e], wy[2],
                                       fmaf(dr[1][j][e], wy[1], dr[0][j][e] * wy[0])));
            };
            // ---- branch channels -> V = f16(elu(t3 + b3a) + b3b): the A fragments of branch_conv3 ----
            uint32_t vf[KS][4];
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                float o[2][4];
                interp_pair(16 * s, o);
                vf[s][0] = act3(o[0][0], o[0][1]);
                vf[s][1] = act3(o[0][2], o[0][3]);
                vf[s][2] = act3(o[1][0], o[1][1]);
                vf[s][3] = act3(o[1][2], o[1][3]);
            }
            float* o0 = a.out + (((size_t)img * Ho + oy) * Wo + ox0 + 16 * seg + g) * CO;
            float* o1 = o0 + 8 * CO;
            if constexpr (NTO == 1) {
                // c_out = 8: one skip n-tile (ldmatrix.x2: its two k halves), W3 fragments by 32-bit reads
                float dr[4][4];
                const uint32_t wb2 = sbase + (uint32_t)((wy0 * U3_RX + seg * 8) * PIXB) +
                                     (uint32_t)((((lane >> 3) & 1) * 8 + (lane & 7)) * PIXB);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    dr[r][0] = dr[r][1] = dr[r][2] = dr[r][3] = 0.f;
                    uint32_t h0, h1, l0, l1;
                    ldmatrix_x2_trans(h0, h1, wb2 + Cfg::OFF_HI + (uint32_t)(r * U3_RX * PIXB + CB * 2));
                    ldmatrix_x2_trans(l0, l1, wb2 + Cfg::OFF_LO + (uint32_t)(r * U3_RX * PIXB + CB * 2));
                    mma_16816(dr[r], hx, l0, l1);
                    mma_16816(dr[r], hx, h0, h1);
                }
                float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    const uint8_t* row = smem + Cfg::OFF_W3 + g * Cfg::WP + (16 * s + 2 * t) * 2;
                    mma_16816(d, vf[s], *reinterpret_cast<const uint32_t*>(row),
                              *reinterpret_cast<const uint32_t*>(row + 16));
                }
                float k[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    k[e] = fmaf(dr[3][e], wy[3], fmaf(dr[2][e], wy[2], fmaf(dr[1][e], wy[1], dr[0][e] * wy[0])));
                *reinterpret_cast<float2*>(o0 + 2 * t) = make_float2(d[0] + k[0] + a.bsum, d[1] + k[1] + a.bsum);
                *reinterpret_cast<float2*>(o1 + 2 * t) = make_float2(d[2] + k[2] + a.bsum, d[3] + k[3] + a.bsum);
            } else {
#pragma unroll
                for (int p = 0; p < NTO / 2; ++p) {
                    float d[2][4];
#pragma unroll
                    for (int j = 0; j < 2; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        uint32_t bf[4];
                        ldmatrix_x4(bf, sbase + Cfg::OFF_W3 + lo3 + (uint32_t)(16 * p) * Cfg::WP + s * 32);
                        mma_16816(d[0], vf[s], bf[0], bf[1]);
                        mma_16816(d[1], vf[s], bf[2], bf[3]);
                    }
                    float k[2][4];
                    interp_pair(CB + 16 * p, k);
                    const float (&e)[4] = d[0], (&f)[4] = d[1];
                    *reinterpret_cast<float4*>(o0 + 16 * p + 4 * t) =
                        make_float4(e[0] + k[0][0] + a.bsum, e[1] + k[0][1] + a.bsum, f[0] + k[1][0] + a.bsum, f[1] + k[1][1] + a.bsum);
                    *reinterpret_cast<float4*>(o1 + 16 * p + 4 * t) =
                        make_float4(e[2] + k[0][2] + a.bsum, e[3] + k[0][3] + a.bsum, f[2] + k[1][2] + a.bsum, f[3] + k[1][3] + a.bsum);
                }
            }
        }
    }
}

template <int CI>
int launch_up_head(UhArgs a, int sm_count, cudaStream_t stream) {
    using Cfg = UhCfg<CI>;
    auto kern = up_head_mma_kernel<CI>;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set.cur() = true;
    }
    const int64_t ctas = (a.n_mtiles + MU_WARPS - 1) / MU_WARPS;
    const int cap = sm_count * Cfg::MIN_CTAS;
    kern<<<(unsigned)(ctas < cap ? ctas : cap), MU_THREADS, Cfg::SMEM, stream>>>(a);
    return check_launch();
}

template <int CB, int TY>
int launch_up_tail(UtArgs a, int64_t B, int sm_count, cudaStream_t stream) {
    using Cfg = U3Cfg<CB, TY>;
    auto kern = up_tail3_mma_kernel<CB, TY>;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set.cur() = true;
    }
    a.tiles_x = (2 * a.W) / U3_TX;
    a.tiles_per_img = ((2 * a.H) / TY) * a.tiles_x;
    const int64_t n = B * a.tiles_per_img;
    if (n > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    a.n_tiles = (int)n;
    a.fd_img = make_fastdiv(a.tiles_per_img);
    a.fd_tx = make_fastdiv(a.tiles_x);
    const int cap = sm_count * Cfg::MIN_CTAS;
    kern<<<(unsigned)(n < cap ? n : cap), MU_THREADS, Cfg::SMEM, stream>>>(a);
    return check_launch();
}

}  // namespace

bool up_block_mma_supported(int H, int W, int CI) {
    // low-res extent H x W; the tail tiles the 2H x 2W output in 8 x 32 pixel tiles
    return (CI == 16 || CI == 32 || CI == 64 || CI == 128) && H >= 4 && W >= 16 && H % 4 == 0 && W % 16 == 0;
}

size_t up_block_mma_scratch_bytes(int64_t B, int H, int W, int CI) {
    const size_t px = (size_t)B * H * W;
    return (px * CI + px * (CI / 2)) * sizeof(float) + 512;
}

// scalars10: b1a b1b b2a b2b b3a b3b b1c (b4 + b1d)
int up_block_mma(const float* x, float* out, const void* w_packed, const float* scalars8,
                 void* scratch, size_t scratch_bytes, int64_t B, int H, int W, int CI, int sm_count,
                 cudaStream_t stream) {
    if (!x || !out || !w_packed || !scalars8 || !scratch || B <= 0) return VQAE_ERR_BAD_ARG;
    if (!up_block_mma_supported(H, W, CI)) return VQAE_ERR_UNSUPPORTED;
    if (scratch_bytes < up_block_mma_scratch_bytes(B, H, W, CI)) return VQAE_ERR_SCRATCH;
    const int64_t P = B * H * W;
    if ((P + 15) / 16 > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    const int CB = CI, CO = CI / 2;
    float* t2 = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
    float* s = t2 + (size_t)P * CB;
    const __half* w = reinterpret_cast<const __half*>(w_packed);
    UhArgs h;
    h.x = x; h.t2 = t2; h.s = s; h.w = w; h.P = P; h.n_mtiles = (int)((P + 15) / 16);
    h.b1a = scalars8[0]; h.b1b = scalars8[1]; h.b2a = scalars8[2]; h.b2b = scalars8[3]; h.b1c = scalars8[6];
    int rc;
    switch (CI) {
        case 16: rc = launch_up_head<16>(h, sm_count, stream); break;
        case 32: rc = launch_up_head<32>(h, sm_count, stream); break;
        case 64: rc = launch_up_head<64>(h, sm_count, stream); break;
        default: rc = launch_up_head<128>(h, sm_count, stream); break;
    }
    if (rc) return rc;
    UtArgs u;
    u.t2 = t2; u.s = s; u.w3 = w + (size_t)(2 * CB + CO) * CI; u.out = out; u.H = H; u.W = W;
    u.b3a = scalars8[4]; u.b3b = scalars8[5]; u.bsum = scalars8[7];
    switch (CI) {
        case 16: return (2 * H) % 16 == 0 ? launch_up_tail<16, 16>(u, B, sm_count, stream)
                                          : launch_up_tail<16, 8>(u, B, sm_count, stream);
        case 32: return (2 * H) % 16 == 0 ? launch_up_tail<32, 16>(u, B, sm_count, stream)
                                          : launch_up_tail<32, 8>(u, B, sm_count, stream);
        case 64: return launch_up_tail<64, 8>(u, B, sm_count, stream);
        default: return launch_up_tail<128, 8>(u, B, sm_count, stream);
    }
}

}  // namespace vqae
