// PreActFixupResBlock 'up' (c_in -> c_in / 2, x2 bicubic; layers/conv_block.py:196-216 with
// ResizeConv2D, layers/conv.py:4-11) on warp-level tensor-core MMAs, in two kernels.  A 1x1 conv
// without bias commutes with the (linear) upsample, so both ResizeConv2Ds run their conv at LOW
// resolution and only the interpolation + branch_conv3 run at high resolution:
//
//   head (low resolution, pointwise, register-resident like mma_down.cu)
//       A1 = f16(elu(x + b1a) + b1b);  D1 = A1 . W1^T;  U = f16(elu(D1 + b2a) + b2b)
//       T2 = U . W2^T                       -> fp32 [B,H,W,CB]      (branch, before the upsample)
//       S  = f16(x + b1c) . Ws^T            -> fp32 [B,H,W,CO]      (skip, before the upsample)
//   tail (high resolution, CTA = 8 x 32 output pixels)
//       low-res window of T2 | S (8 x 20 pixels, index-clamped like nn.Upsample(bicubic,
//       align_corners=False)) staged in shared memory with cp.async
//       per warp and 16-pixel row segment: vertical 4-tap pass into a warp-private buffer, then the
//       horizontal 4-tap pass straight into A-fragment registers (branch) / accumulator-fragment
//       registers (skip); cubic taps for scale 2: (-9, 67, 225, -27) / 256 and its mirror
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
constexpr int UT2_TY = 8, UT2_TX = 32;                       // output pixels per CTA
constexpr int UT2_RY = UT2_TY / 2 + 4, UT2_RX = UT2_TX / 2 + 4;   // low-res window 8 x 20
constexpr int UT2_VC = 12;                                    // low-res columns of one 16-pixel segment (11 used)

template <int CB>
struct UtCfg {
    static constexpr int CO = CB / 2, CT = CB + CO;           // channels of the staged T2 | S pixel
    static constexpr int KS = CB / 16, NTO = CO / 8;
    static constexpr int PW = CT * 4 + 16;                    // window pixel pitch (bytes), 16 * odd
    static constexpr int PV = CT * 4 + 16;                    // vertical-pass pixel pitch
    static constexpr int WP = CB * 2 + 16;                    // W3 row pitch
    static constexpr uint32_t OFF_WIN = 0;
    static constexpr uint32_t OFF_V = OFF_WIN + UT2_RY * UT2_RX * PW;       // per warp: 12 pixels
    static constexpr uint32_t OFF_W3 = OFF_V + MU_WARPS * UT2_VC * PV;
    static constexpr uint32_t SMEM = OFF_W3 + CO * WP;
    static constexpr int MIN_CTAS = CB <= 32 ? 3 : (CB == 64 ? 2 : 1);
};

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

template <int CB>
__global__ void __launch_bounds__(MU_THREADS, UtCfg<CB>::MIN_CTAS)
up_tail_mma_kernel(UtArgs a) {
    using Cfg = UtCfg<CB>;
    constexpr int CO = Cfg::CO, CT = Cfg::CT, KS = Cfg::KS, NTO = Cfg::NTO, PW = Cfg::PW, PV = Cfg::PV;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = tc::smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.w3);
        constexpr int PP = CB / 8;
        for (int i = tid; i < CO * PP; i += MU_THREADS)
            *reinterpret_cast<uint4*>(smem + Cfg::OFF_W3 + (i / PP) * Cfg::WP + (i % PP) * 16) = __ldg(src + i);
    }
    const ActC act3(a.b3a, a.b3b);
    const uint32_t lo = (uint32_t)((lane & 7) + (lane >> 4) * 8) * Cfg::WP + ((lane >> 3) & 1) * 16;
    uint8_t* vbuf = smem + Cfg::OFF_V + warp * UT2_VC * PV;
    const int Ho = 2 * a.H, Wo = 2 * a.W;
    constexpr float WE[4] = {-9.f / 256.f, 67.f / 256.f, 225.f / 256.f, -27.f / 256.f};    // even index
    constexpr float WO[4] = {-27.f / 256.f, 225.f / 256.f, 67.f / 256.f, -9.f / 256.f};    // odd index

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int img = a.fd_img.d == 1 ? tile : a.fd_img.div(tile);
        const int rem = tile - img * a.tiles_per_img;
        const int ty = a.fd_tx.d == 1 ? rem : a.fd_tx.div(rem);
        const int oy0 = ty * UT2_TY, ox0 = (rem - ty * a.tiles_x) * UT2_TX;
        const int ly0 = oy0 / 2 - 2, lx0 = ox0 / 2 - 2;          // low-res origin of the window
        const float* t2 = a.t2 + (size_t)img * a.H * a.W * CB;
        const float* sk = a.s + (size_t)img * a.H * a.W * CO;

        __syncthreads();                                          // previous tile's window reads done
        // ---- stage the T2 | S window, source indices clamped at the image border ----
        constexpr int PPP = CT / 4;                               // 16-byte pieces per pixel
        for (int i = tid; i < UT2_RY * UT2_RX * PPP; i += MU_THREADS) {
            const int p = i / PPP, piece = i - p * PPP;
            const int ry = p / UT2_RX, rx = p - ry * UT2_RX;
            const int y = min(max(ly0 + ry, 0), a.H - 1), x = min(max(lx0 + rx, 0), a.W - 1);
            const float* src = piece < CB / 4 ? t2 + ((size_t)y * a.W + x) * CB + 4 * piece
                                              : sk + ((size_t)y * a.W + x) * CO + 4 * (piece - CB / 4);
            cp_async16(sbase + Cfg::OFF_WIN + p * PW + piece * 16, src);
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_all;" ::: "memory");
        __syncthreads();

        // ---- 16 M-tiles (8 rows x 2 segments), two per warp ----
#pragma unroll 1
        for (int mt = warp; mt < 2 * UT2_TY; mt += MU_WARPS) {
            const int ry_o = mt >> 1, seg = mt & 1;
            const int oy = oy0 + ry_o;
            // vertical taps of this output row (warp-uniform); window-local first source row
            const int wy0 = (oy >> 1) - 2 + (oy & 1) - ly0;
            float wy[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) wy[k] = (oy & 1) ? WO[k] : WE[k];
            // vertical pass: window columns c_base .. c_base + 11 of every channel -> vbuf
            const int c_base = seg * 8;                           // (ox0 + 16 seg) / 2 - 2 - lx0
            __syncwarp();
            for (int i = lane; i < UT2_VC * PPP; i += 32) {
                const int col = i / PPP, piece = i - col * PPP;
                const uint8_t* src = smem + Cfg::OFF_WIN + ((wy0 * UT2_RX) + c_base + col) * PW + piece * 16;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 v = *reinterpret_cast<const float4*>(src + k * UT2_RX * PW);
                    acc.x = fmaf(v.x, wy[k], acc.x); acc.y = fmaf(v.y, wy[k], acc.y);
                    acc.z = fmaf(v.z, wy[k], acc.z); acc.w = fmaf(v.w, wy[k], acc.w);
                }
                *reinterpret_cast<float4*>(vbuf + col * PV + piece * 16) = acc;
            }
            __syncwarp();
            // horizontal pass for this lane's two pixels (fragment rows g and g + 8): output column
            // ox = ox0 + 16 seg + g (+ 8); its four source columns start at (ox >> 1) - 2 + (ox & 1),
            // i.e. vbuf column (g >> 1) + (g & 1) (+ 4 for row g + 8), taps by parity of g
            const int vc0 = (g >> 1) + (g & 1);
            float wx[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) wx[k] = (g & 1) ? WO[k] : WE[k];
            auto hpass = [&](int vcol, int ch_byte) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 v = *reinterpret_cast<const float4*>(vbuf + (vcol + k) * PV + ch_byte);
                    acc.x = fmaf(v.x, wx[k], acc.x); acc.y = fmaf(v.y, wx[k], acc.y);
                    acc.z = fmaf(v.z, wx[k], acc.z); acc.w = fmaf(v.w, wx[k], acc.w);
                }
                return acc;
            };
            uint32_t vf[KS][4];
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const float4 v0 = hpass(vc0, (16 * s + 4 * t) * 4);
                const float4 v1 = hpass(vc0 + 4, (16 * s + 4 * t) * 4);
                vf[s][0] = act3(v0.x, v0.y);
                vf[s][1] = act3(v1.x, v1.y);
                vf[s][2] = act3(v0.z, v0.w);
                vf[s][3] = act3(v1.z, v1.w);
            }
            float d[NTO][4];
#pragma unroll
            for (int j = 0; j < NTO; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
            float* o0 = a.out + (((size_t)img * Ho + oy) * Wo + ox0 + 16 * seg + g) * CO;
            float* o1 = o0 + 8 * CO;
            if constexpr (NTO == 1) {
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    const uint8_t* row = smem + Cfg::OFF_W3 + g * Cfg::WP + (16 * s + 2 * t) * 2;
                    mma_16816(d[0], vf[s], *reinterpret_cast<const uint32_t*>(row),
                              *reinterpret_cast<const uint32_t*>(row + 16));
                }
                // skip channels 2t, 2t + 1 of both pixels
                auto hpass2 = [&](int vcol) {
                    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 v = *reinterpret_cast<const float2*>(vbuf + (vcol + k) * PV + (CB + 2 * t) * 4);
                        acc.x = fmaf(v.x, wx[k], acc.x); acc.y = fmaf(v.y, wx[k], acc.y);
                    }
                    return acc;
                };
                const float2 k0 = hpass2(vc0), k1 = hpass2(vc0 + 4);
                *reinterpret_cast<float2*>(o0 + 2 * t) = make_float2(d[0][0] + k0.x + a.bsum, d[0][1] + k0.y + a.bsum);
                *reinterpret_cast<float2*>(o1 + 2 * t) = make_float2(d[0][2] + k1.x + a.bsum, d[0][3] + k1.y + a.bsum);
            } else {
#pragma unroll
                for (int p = 0; p < NTO / 2; ++p)
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        uint32_t bf[4];
                        ldmatrix_x4(bf, sbase + Cfg::OFF_W3 + lo + (uint32_t)(16 * p) * Cfg::WP + s * 32);
                        mma_16816(d[2 * p], vf[s], bf[0], bf[1]);
                        mma_16816(d[2 * p + 1], vf[s], bf[2], bf[3]);
                    }
#pragma unroll
                for (int p = 0; p < NTO / 2; ++p) {
                    const float4 k0 = hpass(vc0, (CB + 16 * p + 4 * t) * 4);
                    const float4 k1 = hpass(vc0 + 4, (CB + 16 * p + 4 * t) * 4);
                    const float (&e)[4] = d[2 * p], (&f)[4] = d[2 * p + 1];
                    *reinterpret_cast<float4*>(o0 + 16 * p + 4 * t) =
                        make_float4(e[0] + k0.x + a.bsum, e[1] + k0.y + a.bsum, f[0] + k0.z + a.bsum, f[1] + k0.w + a.bsum);
                    *reinterpret_cast<float4*>(o1 + 16 * p + 4 * t) =
                        make_float4(e[2] + k1.x + a.bsum, e[3] + k1.y + a.bsum, f[2] + k1.z + a.bsum, f[3] + k1.w + a.bsum);
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

template <int CB>
int launch_up_tail(UtArgs a, int64_t B, int sm_count, cudaStream_t stream) {
    using Cfg = UtCfg<CB>;
    auto kern = up_tail_mma_kernel<CB>;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set.cur() = true;
    }
    a.tiles_x = (2 * a.W) / UT2_TX;
    a.tiles_per_img = ((2 * a.H) / UT2_TY) * a.tiles_x;
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
        case 16: return launch_up_tail<16>(u, B, sm_count, stream);
        case 32: return launch_up_tail<32>(u, B, sm_count, stream);
        case 64: return launch_up_tail<64>(u, B, sm_count, stream);
        default: return launch_up_tail<128>(u, B, sm_count, stream);
    }
}

}  // namespace vqae
