// MBConv (layers/conv_block.py:240-321; scope row f-4), the alternative EfficientNetV2-style block the
// reference can build its pyramids and trunks from (conf/model/encoder/efficientnetv2.yaml), in eval mode:
//
//   t1   = SiLU(BN1(W1 . x))                         1x1 expand  C_in -> C_mid = expand_ratio * max(C_in, C_out)
//   t2   = SiLU(BN2(depthwise(t1)))                  3x3 circular ('same'), 2x2 stride 2 ('down'),
//                                                    2x2 stride-2 transposed ('up'), one filter per channel
//   gate = sigmoid(FC2 . SiLU(FC1 . mean_hw(t2)))    SELayer (layers/misc.py:7-30)
//   out  = BN3(W3 . (t2 * gate)) + skip(x)           skip = x, or a full 1x1 / 2x2 s2 / transposed 2x2 s2 conv
//
// BatchNorm in eval mode is a per-channel affine map, folded by the host into (scale, shift).
// NHWC fp32 throughout; four kernels per block:
//   pointwise_gemm_kernel   every full conv on the path is a GEMM over pixels ([P, K] x [N, K]^T) with a
//                           mode-dependent A gather (plain rows, 2x2 space-to-depth, transposed-conv parity
//                           classes), an optional per-(image, channel) gate on A, and a fused epilogue
//                           (affine, SiLU, residual).  64 x 64 x 16 shared-memory tiles, 4 x 4 per thread.
//   depthwise_kernel        one thread owns 4 channels of one output row: conv + affine + SiLU + the row's
//                           contribution to the squeeze (a fixed summation order, no atomics)
//   se_gate_kernel          per image: mean -> FC -> SiLU -> FC -> sigmoid
// The arithmetic is plain fp32 FFMA: this block is API generality of the reference that its shipped
// configuration does not use; it is built for parity, not tuned.
#include "common.cuh"
#include "kernels.cuh"

namespace vqae {
namespace {

__device__ __forceinline__ float silu(float v) { return v / (1.0f + expf(-v)); }

enum { PW_DIRECT = 0, PW_S2D = 1, PW_CONVT = 2 };

struct PwArgs {
    const float* a;       // input activations, NHWC [B, Hi, Wi, Cin]
    const float* w;       // [N][K] row-major; PW_CONVT: [4][N][K], one matrix per output parity (dy, dx)
    const float* scale;   // [N] or null (1)
    const float* shift;   // [N] or null (0)
    const float* gate;    // [B][K] or null: A[p, k] *= gate[image(p), k]
    const float* res;     // [P_out][N] or null, added after the activation
    float* out;           // NHWC [B, Ho, Wo, N]
    int64_t P;            // GEMM rows of this launch (PW_CONVT: input pixels, per parity class)
    int K, N, act, mode;
    int Hi, Wi, Cin;      // input geometry
};

constexpr int PW_BM = 64, PW_BN = 64, PW_BK = 16;

__global__ void __launch_bounds__(256)
pointwise_gemm_kernel(PwArgs g) {
    __shared__ __align__(16) float As[PW_BK][PW_BM + 4];
    __shared__ __align__(16) float Ws[PW_BK][PW_BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t p0 = (int64_t)blockIdx.x * PW_BM;
    const int n0 = blockIdx.y * PW_BN;
    const int q = blockIdx.z;                              // PW_CONVT: parity class dy * 2 + dx
    const float* w = g.w + (size_t)q * g.N * g.K;
    const int Ho = g.mode == PW_S2D ? g.Hi / 2 : g.Hi, Wo = g.mode == PW_S2D ? g.Wi / 2 : g.Wi;

    // this thread's A row (pixel) for the loads: row = tid / 4, k-quad = tid % 4
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const int64_t lp = p0 + lrow;
    const bool lvalid = lp < g.P;
    int64_t img = 0;
    int oy = 0, ox = 0;
    if (lvalid) {
        img = lp / ((int64_t)Ho * Wo);
        const int r = (int)(lp - img * Ho * Wo);
        oy = r / Wo;
        ox = r - oy * Wo;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += PW_BK) {
        const int k = k0 + lk;
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lvalid && k < g.K) {
            const float* src;
            if (g.mode == PW_S2D) {                        // k = tap * Cin + ci, tap = dy * 2 + dx
                const int tap = k / g.Cin, ci = k - tap * g.Cin;
                src = g.a + ((img * g.Hi + (2 * oy + (tap >> 1))) * g.Wi + (2 * ox + (tap & 1))) * g.Cin + ci;
            } else {
                src = g.a + lp * g.Cin + k;                // PW_DIRECT / PW_CONVT: K = Cin
            }
            av = __ldg(reinterpret_cast<const float4*>(src));
            if (g.gate) {
                const float4 gv = __ldg(reinterpret_cast<const float4*>(g.gate + img * g.K + k));
                av.x *= gv.x; av.y *= gv.y; av.z *= gv.z; av.w *= gv.w;
            }
        }
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + lrow < g.N && k < g.K)
            wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)(n0 + lrow) * g.K + k));
        __syncthreads();
        As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
        Ws[lk + 0][lrow] = wv.x; Ws[lk + 1][lrow] = wv.y; Ws[lk + 2][lrow] = wv.z; Ws[lk + 3][lrow] = wv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < PW_BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float ar[4] = {a4.x, a4.y, a4.z, a4.w}, wr[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
        }
    }
    const int n = n0 + tx * 4;
    if (n >= g.N) return;
    float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
    if (g.scale) { const float4 v = __ldg(reinterpret_cast<const float4*>(g.scale + n)); sc[0] = v.x; sc[1] = v.y; sc[2] = v.z; sc[3] = v.w; }
    if (g.shift) { const float4 v = __ldg(reinterpret_cast<const float4*>(g.shift + n)); sh[0] = v.x; sh[1] = v.y; sh[2] = v.z; sh[3] = v.w; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t p = p0 + ty * 4 + i;
        if (p >= g.P) continue;
        int64_t po = p;                                     // output pixel index
        if (g.mode == PW_CONVT) {                           // input pixel (y, x) -> output (2y + dy, 2x + dx)
            const int64_t im = p / ((int64_t)g.Hi * g.Wi);
            const int r = (int)(p - im * g.Hi * g.Wi);
            const int y = r / g.Wi, x = r - y * g.Wi;
            po = (im * (2 * g.Hi) + (2 * y + (q >> 1))) * (2 * g.Wi) + (2 * x + (q & 1));
        }
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = fmaf(acc[i][j], sc[j], sh[j]);
            if (g.act) v[j] = silu(v[j]);
        }
        if (g.res) {
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(g.res + po * g.N + n));
            v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
        }
        *reinterpret_cast<float4*>(g.out + po * g.N + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

enum { DW_SAME = 0, DW_DOWN = 1, DW_UP = 2 };

// grid (Ho, B, ceil(C / (4 * 64))), block 64: a thread owns 4 channels of one output row
__global__ void __launch_bounds__(64)
depthwise_kernel(const float* __restrict__ in, const float* __restrict__ w /* [taps][C] */,
                 const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out,
                 float* __restrict__ row_sums /* [B][Ho][C] or null */, int Hi, int Wi, int C, int mode) {
    const int c = (blockIdx.z * 64 + threadIdx.x) * 4;
    if (c >= C) return;
    const int oy = blockIdx.x, b = blockIdx.y;
    const int Ho = mode == DW_DOWN ? Hi / 2 : (mode == DW_UP ? Hi * 2 : Hi);
    const int Wo = mode == DW_DOWN ? Wi / 2 : (mode == DW_UP ? Wi * 2 : Wi);
    const int taps = mode == DW_SAME ? 9 : 4;
    float4 wt[9];
    for (int t = 0; t < taps; ++t) wt[t] = __ldg(reinterpret_cast<const float4*>(w + (size_t)t * C + c));
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale + c));
    if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift + c));
    const float* img = in + (size_t)b * Hi * Wi * C + c;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ox = 0; ox < Wo; ++ox) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        auto tap = [&](int y, int x, const float4& wv) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(img + ((size_t)y * Wi + x) * C));
            a.x = fmaf(v.x, wv.x, a.x); a.y = fmaf(v.y, wv.y, a.y);
            a.z = fmaf(v.z, wv.z, a.z); a.w = fmaf(v.w, wv.w, a.w);
        };
        if (mode == DW_SAME) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int y = (oy + dy - 1 + Hi) % Hi;              // circular padding
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) tap(y, (ox + dx - 1 + Wi) % Wi, wt[dy * 3 + dx]);
            }
        } else if (mode == DW_DOWN) {
#pragma unroll
            for (int t = 0; t < 4; ++t) tap(2 * oy + (t >> 1), 2 * ox + (t & 1), wt[t]);
        } else {                                                    // transposed 2x2 stride 2
            tap(oy >> 1, ox >> 1, wt[(oy & 1) * 2 + (ox & 1)]);
        }
        float4 v;
        v.x = silu(fmaf(a.x, sc.x, sh.x)); v.y = silu(fmaf(a.y, sc.y, sh.y));
        v.z = silu(fmaf(a.z, sc.z, sh.z)); v.w = silu(fmaf(a.w, sc.w, sh.w));
        *reinterpret_cast<float4*>(out + (((size_t)b * Ho + oy) * Wo + ox) * C + c) = v;
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
    if (row_sums) *reinterpret_cast<float4*>(row_sums + ((size_t)b * Ho + oy) * C + c) = sum;
}

// grid B, block 256.  gate[b][c] = sigmoid(b2[c] + sum_j w2[c][j] * SiLU(b1[j] + sum_c' w1[j][c'] * mean[c']))
__global__ void __launch_bounds__(256)
se_gate_kernel(const float* __restrict__ row_sums, int Ho, int C, float inv_hw, const float* __restrict__ w1,
               const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2, int CI,
               float* __restrict__ gate) {
    extern __shared__ float sm[];                                   // mean[C] | hidden[CI]
    float* mean = sm;
    float* hid = sm + C;
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < Ho; ++r) s += row_sums[((size_t)b * Ho + r) * C + c];   // rows in order
        mean[c] = s * inv_hw;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < CI; j += blockDim.x) {
        float s = b1[j];
        for (int c = 0; c < C; ++c) s = fmaf(w1[(size_t)j * C + c], mean[c], s);
        hid[j] = silu(s);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = b2[c];
        for (int j = 0; j < CI; ++j) s = fmaf(w2[(size_t)c * CI + j], hid[j], s);
        gate[(size_t)b * C + c] = 1.0f / (1.0f + expf(-s));
    }
}

}  // namespace
}  // namespace vqae

using namespace vqae;

extern "C" {

int vqae_pointwise_conv_f32(const float* a, const float* w, const float* scale, const float* shift,
                            const float* gate, const float* res, float* out, int64_t batch, int hi, int wi,
                            int c_in, int n_out, int mode, int act_silu, void* stream) {
    if (!a || !w || !out || batch <= 0 || hi <= 0 || wi <= 0 || c_in <= 0 || n_out <= 0) return VQAE_ERR_BAD_ARG;
    if (mode < PW_DIRECT || mode > PW_CONVT) return VQAE_ERR_BAD_ARG;
    if ((c_in & 3) || (n_out & 3)) return VQAE_ERR_UNSUPPORTED;
    if (mode == PW_S2D && ((hi | wi) & 1)) return VQAE_ERR_UNSUPPORTED;
    if (mode != PW_DIRECT && gate) return VQAE_ERR_UNSUPPORTED;
    PwArgs g{};
    g.a = a; g.w = w; g.scale = scale; g.shift = shift; g.gate = gate; g.res = res; g.out = out;
    g.mode = mode; g.act = act_silu; g.Hi = hi; g.Wi = wi; g.Cin = c_in; g.N = n_out;
    g.K = mode == PW_S2D ? 4 * c_in : c_in;
    g.P = mode == PW_S2D ? batch * (hi / 2) * (wi / 2) : batch * hi * wi;
    dim3 grid(ceil_div_u(g.P, PW_BM), ceil_div_u(n_out, PW_BN), mode == PW_CONVT ? 4 : 1);
    pointwise_gemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g);
    return check_launch();
}

int vqae_depthwise_conv_f32(const float* in, const float* w_taps, const float* scale, const float* shift,
                            float* out, float* row_sums, int64_t batch, int hi, int wi, int c, int mode,
                            void* stream) {
    if (!in || !w_taps || !out || batch <= 0 || hi <= 0 || wi <= 0 || c <= 0) return VQAE_ERR_BAD_ARG;
    if (mode < DW_SAME || mode > DW_UP) return VQAE_ERR_BAD_ARG;
    if (c & 3) return VQAE_ERR_UNSUPPORTED;
    if (mode == DW_DOWN && ((hi | wi) & 1)) return VQAE_ERR_UNSUPPORTED;
    if (batch > 65535) return VQAE_ERR_UNSUPPORTED;
    const int ho = mode == DW_DOWN ? hi / 2 : (mode == DW_UP ? hi * 2 : hi);
    dim3 grid(ho, (unsigned)batch, ceil_div_u(c, 256));
    depthwise_kernel<<<grid, 64, 0, (cudaStream_t)stream>>>(in, w_taps, scale, shift, out, row_sums, hi, wi, c, mode);
    return check_launch();
}

int vqae_se_gate_f32(const float* row_sums, int64_t batch, int rows, int pixels_per_image, int c,
                     const float* w1, const float* b1, const float* w2, const float* b2, int c_hidden,
                     float* gate, void* stream) {
    if (!row_sums || !w1 || !b1 || !w2 || !b2 || !gate || batch <= 0 || rows <= 0 || pixels_per_image <= 0 ||
        c <= 0 || c_hidden <= 0)
        return VQAE_ERR_BAD_ARG;
    const size_t smem = (size_t)(c + c_hidden) * sizeof(float);
    if (smem > 48 * 1024) return VQAE_ERR_UNSUPPORTED;
    se_gate_kernel<<<(unsigned)batch, 256, smem, (cudaStream_t)stream>>>(
        row_sums, rows, c, 1.0f / (float)pixels_per_image, w1, b1, w2, b2, c_hidden, gate);
    return check_launch();
}

}  // extern "C"
