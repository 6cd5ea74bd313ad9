// PreActFixupResBlock 'down' (c_in -> 2 c_in, stride 2; layers/conv_block.py:196-216 with the conv
// specs of pre_activation_fixup.yaml:35-45) on warp-level tensor-core MMAs, every intermediate in
// REGISTERS: a 2x2 stride-2 conv has no halo, so a warp carries 16 output pixels from the four input
// pixels under each of them to the output without touching shared memory for activations.
//
//   for each of the four positions (dy, dx) of the 2x2 window:
//       x      = input pixel (2i + dy, 2j + dx)                 128-bit loads into the A-fragment slots
//       A1     = f16(elu(x + b1a) + b1b),  As = f16(x + b1c)
//       D1     = A1 . W1^T                                      (c_in -> c_out)
//       U      = f16(elu(D1 + b2a) + b2b)                       accumulator fragments = next A fragments
//       D2    += U . W2[dy,dx]^T                                the 2x2 conv: four taps = four GEMMs
//       D3    += As . Ws[dy,dx]^T                               the skip conv (same window)
//   V   = f16(elu(D2 + b3a) + b3b);  D3 += V . (scale W3)^T
//   out = D3 + (b4 + b1d)                                        128-bit stores
//
// SPLIT = true is the fp32-accurate form ("fp32tc" precision, see tc_split.cu): every operand is a
// pair hi + lo of fp16 numbers, every product three MMAs (hi.hi + lo.hi + hi.lo), the weights are
// pre-multiplied by powers of two at pack time (the accumulators by the inverse) and the activation
// is the fp32-grade elu1_tc (common.cuh).
//
// Weights sit in shared memory ([n][k] rows, padded pitch) and are read as B fragments with ldmatrix.
// As in mma_same.cu the input-channel order of W1 / Ws and the output-channel order of W3 / Ws are
// permuted at pack time so that a lane's fragment slots are four consecutive channels in memory.
#include "common.cuh"
#include "kernels.cuh"
#include "mma_common.cuh"

namespace vqae {
namespace {

using namespace mma;

constexpr int MD_WARPS = 8;
constexpr int MD_THREADS = MD_WARPS * 32;

template <int CI>
struct MdCfg {
    static constexpr int CO = 2 * CI;
    static constexpr bool K8 = (CI == 8);                  // input GEMMs with K = 8 (m16n8k8)
    static constexpr int KSI = K8 ? 1 : CI / 16;           // k-steps of the input GEMMs
    static constexpr int KSO = CO / 16;                    // k-steps of the CO x CO GEMMs
    static constexpr int NT = CO / 8;                      // n-tiles
    static constexpr int WPI = K8 ? 16 : CI * 2 + 16;      // row pitch of [CO][CI] matrices (bytes)
    static constexpr int WPO = CO * 2 + 16;                // row pitch of [CO][CO] matrices
    static constexpr uint32_t OFF_W1 = 0;
    static constexpr uint32_t OFF_W2 = OFF_W1 + CO * WPI;           // 4 taps
    static constexpr uint32_t OFF_W3 = OFF_W2 + 4 * CO * WPO;
    static constexpr uint32_t OFF_WS = OFF_W3 + CO * WPO;           // 4 taps
    static constexpr uint32_t SMEM = OFF_WS + 4 * CO * WPI;              // one set (hi); SPLIT: x 2
    static constexpr int MIN_CTAS = CI == 8 ? 4 : (CI == 16 ? 2 : 1);
    // elements of the packed global weights (same order, dense [n][k])
    static constexpr int N_W1 = CO * CI, N_W2 = 4 * CO * CO, N_W3 = CO * CO, N_WS = 4 * CO * CI;
};

struct MdArgs {
    const float* x;               // NHWC fp32 [B,H,W,CI]
    float* out;                   // NHWC fp32 [B,H/2,W/2,2CI]
    const __half* w;              // [W1 | W2 x4 | scale*W3 | Ws x4], dense [n][k] fp16 (pack.cu)
    const __half* w_lo;           // SPLIT: the low halves, same layout
    float inv1, inv2, inv3;       // SPLIT: 1 / premul of W1, W2, (W3 and Ws)
    int n_mtiles, H, W, mt_per_row, mt_per_img;
    FastDiv fd_img, fd_row;
    float b1a, b1b, b2a, b2b, b3a, b3b, b1c, bsum;
};

// exact pre-activation of the fp32 path on a scaled accumulator: elu(v * mul + pre) + post
struct ActX {
    float pre, post, mul;
    __device__ __forceinline__ float operator()(float v) const { return elu1_tc(fmaf(v, mul, pre)) + post; }
};
// (f0, f1) -> fp16 pair of the high halves and fp16 pair of the remainders
__device__ __forceinline__ void split2(float f0, float f1, uint32_t& hi, uint32_t& lo) {
    const __half2 hh = __floats2half2_rn(f0, f1);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&hh);
    lo = *reinterpret_cast<const uint32_t*>(&ll);
}

// resident CTAs per SM of the SPLIT form (twice the shared memory, ~1.5x the registers)
template <int CI>
constexpr int md_split_ctas() { return CI == 8 ? 2 : 1; }

template <int CI, bool SPLIT>
__global__ void __launch_bounds__(MD_THREADS, SPLIT ? md_split_ctas<CI>() : MdCfg<CI>::MIN_CTAS)
down_block_mma_kernel(MdArgs a) {
    using Cfg = MdCfg<CI>;
    constexpr int CO = Cfg::CO, NT = Cfg::NT, KSI = Cfg::KSI, KSO = Cfg::KSO;
    constexpr bool K8 = Cfg::K8;
    extern __shared__ __align__(128) uint8_t smem_all[];
    const uint32_t sbase = tc::smem_u32(smem_all);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    // ---- weights -> shared memory, 16-byte pieces, rows re-pitched ----
    constexpr uint32_t LO = Cfg::SMEM;                                // offset of the lo set (SPLIT)
#pragma unroll
    for (int set = 0; set < (SPLIT ? 2 : 1); ++set) {
        const uint4* src = reinterpret_cast<const uint4*>(set ? a.w_lo : a.w);
        uint8_t* const smem = smem_all + set * LO;
        constexpr int PI = CI / 8, PO = CO / 8;                       // pieces per row
        constexpr int R1 = CO, R2 = 4 * CO, R3 = CO, RS = 4 * CO;     // rows per group
        for (int i = tid; i < R1 * PI; i += MD_THREADS)
            *reinterpret_cast<uint4*>(smem + Cfg::OFF_W1 + (i / PI) * Cfg::WPI + (i % PI) * 16) = __ldg(src + i);
        src += R1 * PI;
        for (int i = tid; i < R2 * PO; i += MD_THREADS)
            *reinterpret_cast<uint4*>(smem + Cfg::OFF_W2 + (i / PO) * Cfg::WPO + (i % PO) * 16) = __ldg(src + i);
        src += R2 * PO;
        for (int i = tid; i < R3 * PO; i += MD_THREADS)
            *reinterpret_cast<uint4*>(smem + Cfg::OFF_W3 + (i / PO) * Cfg::WPO + (i % PO) * 16) = __ldg(src + i);
        src += R3 * PO;
        for (int i = tid; i < RS * PI; i += MD_THREADS)
            *reinterpret_cast<uint4*>(smem + Cfg::OFF_WS + (i / PI) * Cfg::WPI + (i % PI) * 16) = __ldg(src + i);
    }
    __syncthreads();
    const ActC act1(a.b1a, a.b1b), act2(a.b2a, a.b2b), act3(a.b3a, a.b3b);
    const ActX ex1{a.b1a, a.b1b, 1.f}, ex2{a.b2a, a.b2b, a.inv1}, ex3{a.b3a, a.b3b, a.inv2};
    const float2 c1c = make_float2(a.b1c, a.b1c);

    // ldmatrix lane addresses of B fragments: matrices (n 0-7 | 8-15) x (k 0-7 | 8-15) of an n-tile pair
    const uint32_t lo_i = (uint32_t)((lane & 7) + (K8 ? ((lane >> 3) & 1) * 8 : (lane >> 4) * 8)) * Cfg::WPI +
                          (K8 ? 0 : ((lane >> 3) & 1) * 16);
    const uint32_t lo_o = (uint32_t)((lane & 7) + (lane >> 4) * 8) * Cfg::WPO + ((lane >> 3) & 1) * 16;

    // D[j] += A . W^T for the [CO][CI] matrices (W1, Ws): n-tile pairs, input k-steps
    auto gemm_in = [&](float (&d)[NT][4], const uint32_t (&af)[KSI][4], const uint32_t (&al)[KSI][4],
                       uint32_t wbase) {
#pragma unroll
        for (int p = 0; p < NT / 2; ++p) {
            if constexpr (K8) {
                uint32_t b0, b1;
                ldmatrix_x2(b0, b1, wbase + lo_i + (uint32_t)(16 * p) * Cfg::WPI);
                if constexpr (SPLIT) {
                    uint32_t l0, l1;
                    ldmatrix_x2(l0, l1, wbase + LO + lo_i + (uint32_t)(16 * p) * Cfg::WPI);
                    mma_1688(d[2 * p], al[0][0], al[0][1], b0);
                    mma_1688(d[2 * p + 1], al[0][0], al[0][1], b1);
                    mma_1688(d[2 * p], af[0][0], af[0][1], l0);
                    mma_1688(d[2 * p + 1], af[0][0], af[0][1], l1);
                }
                mma_1688(d[2 * p], af[0][0], af[0][1], b0);
                mma_1688(d[2 * p + 1], af[0][0], af[0][1], b1);
            } else {
#pragma unroll
                for (int s = 0; s < KSI; ++s) {
                    uint32_t bf[4];
                    ldmatrix_x4(bf, wbase + lo_i + (uint32_t)(16 * p) * Cfg::WPI + s * 32);
                    if constexpr (SPLIT) {
                        uint32_t bl[4];
                        ldmatrix_x4(bl, wbase + LO + lo_i + (uint32_t)(16 * p) * Cfg::WPI + s * 32);
                        mma_16816(d[2 * p], al[s], bf[0], bf[1]);
                        mma_16816(d[2 * p + 1], al[s], bf[2], bf[3]);
                        mma_16816(d[2 * p], af[s], bl[0], bl[1]);
                        mma_16816(d[2 * p + 1], af[s], bl[2], bl[3]);
                    }
                    mma_16816(d[2 * p], af[s], bf[0], bf[1]);
                    mma_16816(d[2 * p + 1], af[s], bf[2], bf[3]);
                }
            }
        }
    };
    // D[j] += A . W^T for the [CO][CO] matrices (W2 taps, W3)
    auto gemm_out = [&](float (&d)[NT][4], const uint32_t (&af)[KSO][4], const uint32_t (&al)[KSO][4],
                        uint32_t wbase) {
#pragma unroll
        for (int p = 0; p < NT / 2; ++p)
#pragma unroll
            for (int s = 0; s < KSO; ++s) {
                uint32_t bf[4];
                ldmatrix_x4(bf, wbase + lo_o + (uint32_t)(16 * p) * Cfg::WPO + s * 32);
                if constexpr (SPLIT) {
                    uint32_t bl[4];
                    ldmatrix_x4(bl, wbase + LO + lo_o + (uint32_t)(16 * p) * Cfg::WPO + s * 32);
                    mma_16816(d[2 * p], al[s], bf[0], bf[1]);
                    mma_16816(d[2 * p + 1], al[s], bf[2], bf[3]);
                    mma_16816(d[2 * p], af[s], bl[0], bl[1]);
                    mma_16816(d[2 * p + 1], af[s], bl[2], bl[3]);
                }
                mma_16816(d[2 * p], af[s], bf[0], bf[1]);
                mma_16816(d[2 * p + 1], af[s], bf[2], bf[3]);
            }
    };

    const int Wo = a.W / 2;
    const size_t in_img = (size_t)a.H * a.W * CI, out_img = (size_t)(a.H / 2) * Wo * CO;
    constexpr int XV = K8 ? 1 : 2 * KSI;                  // 128-bit (K8: 64-bit) loads per row pair...

    for (int mt = blockIdx.x * MD_WARPS + warp; mt < a.n_mtiles; mt += gridDim.x * MD_WARPS) {
        // M-tile -> (image, output row, first output column)
        const int img = a.fd_img.d == 1 ? mt : a.fd_img.div(mt);
        const int rem = mt - img * a.mt_per_img;
        const int orow = a.fd_row.d == 1 ? rem : a.fd_row.div(rem);
        const int ocol = (rem - orow * a.mt_per_row) * 16;
        const float* ximg = a.x + img * in_img;
        // input pixel of fragment row g (and g + 8) at position (0, 0): (2 orow, 2 (ocol + g))
        const float* px0 = ximg + ((size_t)(2 * orow) * a.W + 2 * (ocol + g)) * CI + (K8 ? 2 : 4) * t;
        const float* px1 = px0 + 16 * CI;                 // row g + 8: 8 output = 16 input pixels further

        float d2[NT][4], d3[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) d2[j][e] = d3[j][e] = 0.f;

        // x of one position into registers: [row g | row g + 8] x k-steps
        float4 xv[2][2][KSI];                             // [buffer][row][k-step]
        auto load_pos = [&](int pos, int buf) {
            const size_t off = ((size_t)(pos >> 1) * a.W + (pos & 1)) * CI;
#pragma unroll
            for (int s = 0; s < KSI; ++s) {
                if constexpr (K8) {
                    const float2 u0 = __ldg(reinterpret_cast<const float2*>(px0 + off));
                    const float2 u1 = __ldg(reinterpret_cast<const float2*>(px1 + off));
                    xv[buf][0][s] = make_float4(u0.x, u0.y, 0.f, 0.f);
                    xv[buf][1][s] = make_float4(u1.x, u1.y, 0.f, 0.f);
                } else {
                    xv[buf][0][s] = __ldg(reinterpret_cast<const float4*>(px0 + off + 16 * s));
                    xv[buf][1][s] = __ldg(reinterpret_cast<const float4*>(px1 + off + 16 * s));
                }
            }
        };
        load_pos(0, 0);
#pragma unroll
        for (int pos = 0; pos < 4; ++pos) {
            const int buf = pos & 1;
            if (pos + 1 < 4) load_pos(pos + 1, buf ^ 1);
            uint32_t a1[KSI][4], as[KSI][4], a1l[KSI][4], asl[KSI][4];
#pragma unroll
            for (int s = 0; s < KSI; ++s) {
                const float4 v0 = xv[buf][0][s], v1 = xv[buf][1][s];
                if constexpr (SPLIT) {
                    split2(ex1(v0.x), ex1(v0.y), a1[s][0], a1l[s][0]);
                    split2(ex1(v1.x), ex1(v1.y), a1[s][1], a1l[s][1]);
                    split2(v0.x + a.b1c, v0.y + a.b1c, as[s][0], asl[s][0]);
                    split2(v1.x + a.b1c, v1.y + a.b1c, as[s][1], asl[s][1]);
                    if constexpr (!K8) {
                        split2(ex1(v0.z), ex1(v0.w), a1[s][2], a1l[s][2]);
                        split2(ex1(v1.z), ex1(v1.w), a1[s][3], a1l[s][3]);
                        split2(v0.z + a.b1c, v0.w + a.b1c, as[s][2], asl[s][2]);
                        split2(v1.z + a.b1c, v1.w + a.b1c, as[s][3], asl[s][3]);
                    }
                    continue;
                }
                a1[s][0] = act1(v0.x, v0.y);
                a1[s][1] = act1(v1.x, v1.y);
                const float2 s0 = __fadd2_rn(make_float2(v0.x, v0.y), c1c);
                const float2 s1 = __fadd2_rn(make_float2(v1.x, v1.y), c1c);
                as[s][0] = pack_h2(s0.x, s0.y);
                as[s][1] = pack_h2(s1.x, s1.y);
                if constexpr (!K8) {
                    a1[s][2] = act1(v0.z, v0.w);
                    a1[s][3] = act1(v1.z, v1.w);
                    const float2 s2 = __fadd2_rn(make_float2(v0.z, v0.w), c1c);
                    const float2 s3 = __fadd2_rn(make_float2(v1.z, v1.w), c1c);
                    as[s][2] = pack_h2(s2.x, s2.y);
                    as[s][3] = pack_h2(s3.x, s3.y);
                }
            }
            float d1[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j) d1[j][0] = d1[j][1] = d1[j][2] = d1[j][3] = 0.f;
            gemm_in(d1, a1, a1l, sbase + Cfg::OFF_W1);
            uint32_t uf[KSO][4], ul[KSO][4];
#pragma unroll
            for (int s = 0; s < KSO; ++s) {
                if constexpr (SPLIT) {
                    split2(ex2(d1[2 * s][0]), ex2(d1[2 * s][1]), uf[s][0], ul[s][0]);
                    split2(ex2(d1[2 * s][2]), ex2(d1[2 * s][3]), uf[s][1], ul[s][1]);
                    split2(ex2(d1[2 * s + 1][0]), ex2(d1[2 * s + 1][1]), uf[s][2], ul[s][2]);
                    split2(ex2(d1[2 * s + 1][2]), ex2(d1[2 * s + 1][3]), uf[s][3], ul[s][3]);
                    continue;
                }
                uf[s][0] = act2(d1[2 * s][0], d1[2 * s][1]);
                uf[s][1] = act2(d1[2 * s][2], d1[2 * s][3]);
                uf[s][2] = act2(d1[2 * s + 1][0], d1[2 * s + 1][1]);
                uf[s][3] = act2(d1[2 * s + 1][2], d1[2 * s + 1][3]);
            }
            gemm_out(d2, uf, ul, sbase + Cfg::OFF_W2 + (uint32_t)(pos * CO) * Cfg::WPO);
            gemm_in(d3, as, asl, sbase + Cfg::OFF_WS + (uint32_t)(pos * CO) * Cfg::WPI);
        }
        {
            uint32_t vf[KSO][4], vl[KSO][4];
#pragma unroll
            for (int s = 0; s < KSO; ++s) {
                if constexpr (SPLIT) {
                    split2(ex3(d2[2 * s][0]), ex3(d2[2 * s][1]), vf[s][0], vl[s][0]);
                    split2(ex3(d2[2 * s][2]), ex3(d2[2 * s][3]), vf[s][1], vl[s][1]);
                    split2(ex3(d2[2 * s + 1][0]), ex3(d2[2 * s + 1][1]), vf[s][2], vl[s][2]);
                    split2(ex3(d2[2 * s + 1][2]), ex3(d2[2 * s + 1][3]), vf[s][3], vl[s][3]);
                    continue;
                }
                vf[s][0] = act3(d2[2 * s][0], d2[2 * s][1]);
                vf[s][1] = act3(d2[2 * s][2], d2[2 * s][3]);
                vf[s][2] = act3(d2[2 * s + 1][0], d2[2 * s + 1][1]);
                vf[s][3] = act3(d2[2 * s + 1][2], d2[2 * s + 1][3]);
            }
            gemm_out(d3, vf, vl, sbase + Cfg::OFF_W3);
        }
        if constexpr (SPLIT) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) d3[j][e] *= a.inv3;
        }
        // out rows g, g + 8: lane (g, t) owns channels 16p + 4t .. + 3 of n-tile pair p
        float* o0 = a.out + img * out_img + ((size_t)orow * Wo + ocol + g) * CO + 4 * t;
        float* o1 = o0 + 8 * CO;
#pragma unroll
        for (int p = 0; p < NT / 2; ++p) {
            const float (&e)[4] = d3[2 * p], (&f)[4] = d3[2 * p + 1];
            *reinterpret_cast<float4*>(o0 + 16 * p) =
                make_float4(e[0] + a.bsum, e[1] + a.bsum, f[0] + a.bsum, f[1] + a.bsum);
            *reinterpret_cast<float4*>(o1 + 16 * p) =
                make_float4(e[2] + a.bsum, e[3] + a.bsum, f[2] + a.bsum, f[3] + a.bsum);
        }
    }
    (void)XV;
}

template <int CI, bool SPLIT>
int launch_down_mma(MdArgs a, int64_t B, int sm_count, cudaStream_t stream) {
    using Cfg = MdCfg<CI>;
    auto kern = down_block_mma_kernel<CI, SPLIT>;
    constexpr int SMEM = (SPLIT ? 2 : 1) * (int)Cfg::SMEM;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set.cur() = true;
    }
    a.mt_per_row = (a.W / 2) / 16;
    a.mt_per_img = (a.H / 2) * a.mt_per_row;
    const int64_t n = B * a.mt_per_img;
    if (n > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    a.n_mtiles = (int)n;
    a.fd_img = make_fastdiv(a.mt_per_img);
    a.fd_row = make_fastdiv(a.mt_per_row);
    const int64_t ctas_needed = (n + MD_WARPS - 1) / MD_WARPS;
    const int cap = sm_count * (SPLIT ? md_split_ctas<CI>() : Cfg::MIN_CTAS);
    const int grid = ctas_needed < cap ? (int)ctas_needed : cap;
    kern<<<grid, MD_THREADS, SMEM, stream>>>(a);
    return check_launch();
}

}  // namespace

bool down_block_mma_supported(int H, int W, int CI) {
    return (CI == 8 || CI == 16 || CI == 32) && H >= 2 && W >= 32 && H % 2 == 0 && W % 32 == 0;
}

size_t down_block_mma_pack_elems(int CI) {
    const size_t co = 2 * (size_t)CI;
    return co * CI + 4 * co * co + co * co + 4 * co * CI;
}

int down_block_mma(const float* x, float* out, const void* w_packed, const float* scalars8,
                   int64_t B, int H, int W, int CI, int sm_count, cudaStream_t stream) {
    if (!x || !out || !w_packed || !scalars8 || B <= 0) return VQAE_ERR_BAD_ARG;
    if (!down_block_mma_supported(H, W, CI)) return VQAE_ERR_UNSUPPORTED;
    MdArgs a;
    a.x = x; a.out = out; a.w = reinterpret_cast<const __half*>(w_packed);
    a.H = H; a.W = W;
    a.b1a = scalars8[0]; a.b1b = scalars8[1]; a.b2a = scalars8[2]; a.b2b = scalars8[3];
    a.b3a = scalars8[4]; a.b3b = scalars8[5]; a.b1c = scalars8[6]; a.bsum = scalars8[7];
    a.w_lo = nullptr; a.inv1 = a.inv2 = a.inv3 = 1.f;
    switch (CI) {
        case 8: return launch_down_mma<8, false>(a, B, sm_count, stream);
        case 16: return launch_down_mma<16, false>(a, B, sm_count, stream);
        case 32: return launch_down_mma<32, false>(a, B, sm_count, stream);
    }
    return VQAE_ERR_UNSUPPORTED;
}

int down_block_split(const float* x, float* out, const void* w_hi, const void* w_lo,
                     const float* scalars8, const float* premul3, int64_t B, int H, int W, int CI,
                     int sm_count, cudaStream_t stream) {
    if (!x || !out || !w_hi || !w_lo || !scalars8 || !premul3 || B <= 0) return VQAE_ERR_BAD_ARG;
    if (!down_block_mma_supported(H, W, CI)) return VQAE_ERR_UNSUPPORTED;
    for (int i = 0; i < 3; ++i)
        if (!(premul3[i] > 0.f)) return VQAE_ERR_BAD_ARG;
    MdArgs a;
    a.x = x; a.out = out;
    a.w = reinterpret_cast<const __half*>(w_hi);
    a.w_lo = reinterpret_cast<const __half*>(w_lo);
    a.H = H; a.W = W;
    a.b1a = scalars8[0]; a.b1b = scalars8[1]; a.b2a = scalars8[2]; a.b2b = scalars8[3];
    a.b3a = scalars8[4]; a.b3b = scalars8[5]; a.b1c = scalars8[6]; a.bsum = scalars8[7];
    a.inv1 = 1.f / premul3[0]; a.inv2 = 1.f / premul3[1]; a.inv3 = 1.f / premul3[2];
    switch (CI) {
        case 8: return launch_down_mma<8, true>(a, B, sm_count, stream);
        case 16: return launch_down_mma<16, true>(a, B, sm_count, stream);
        case 32: return launch_down_mma<32, true>(a, B, sm_count, stream);
    }
    return VQAE_ERR_UNSUPPORTED;
}

}  // namespace vqae
