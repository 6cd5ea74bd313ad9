// fp32 CUDA-core implicit-GEMM convolutions for the exact-parity path.
//
// One kernel template covers every branch/skip conv of PreActFixupResBlock
// (reference: vq_ae/layers/conv_block.py:196-216, conv specs in
// conf/model/layers/conv_block/pre_activation_fixup.yaml:34-74):
//   CONV_1x1       proj2d      (branch_conv1/3, and the 1x1 that follows the bicubic upsample)
//   CONV_2x2S2     down2d      (kernel 2, stride 2, no padding)
//   CONV_3x3_CIRC  same2d      (kernel 3, padding 1, padding_mode circular)
// with the Fixup pre-activation  act(x + a) + b  fused into the operand load and
// acc * scale + bias (+ residual)  fused into the epilogue.  Activations are NHWC fp32, weights
// are packed [tap][Cin][Cout].  GEMM view: M = B*Ho*Wo pixels, N = Cout, K = taps*Cin.
#include "common.cuh"
#include "kernels.cuh"

namespace vqae {

namespace {

constexpr int BM = 128;       // pixels per CTA
constexpr int NTHREADS = 256;

struct ConvArgs {
    const float* in;
    const float* w;
    float* out;
    const float* res;
    int64_t M;
    int Hi, Wi, Cin, Ho, Wo, Cout;
    PreOp pre;
    float scale, bias;
};

template <int KIND>
__device__ __forceinline__ int num_taps() {
    return KIND == CONV_1x1 ? 1 : (KIND == CONV_2x2S2 ? 4 : 9);
}

// input pixel (linear NHWC pixel index) feeding output pixel (b, oy, ox) through tap t
template <int KIND>
__device__ __forceinline__ int64_t in_pixel(int b, int oy, int ox, int t, int Hi, int Wi) {
    int iy, ix;
    if (KIND == CONV_1x1) {
        iy = oy;
        ix = ox;
    } else if (KIND == CONV_2x2S2) {
        iy = 2 * oy + (t >> 1);
        ix = 2 * ox + (t & 1);
    } else {
        iy = oy + t / 3 - 1;
        ix = ox + t % 3 - 1;
        iy = iy < 0 ? iy + Hi : (iy >= Hi ? iy - Hi : iy);
        ix = ix < 0 ? ix + Wi : (ix >= Wi ? ix - Wi : ix);
    }
    return ((int64_t)b * Hi + iy) * Wi + ix;
}

template <int KIND, int BN, int BK>
__global__ void __launch_bounds__(NTHREADS) conv_f32_kernel(ConvArgs a) {
    constexpr int NT = BN / 4;            // threads along N, 4 channels each
    constexpr int MT = NTHREADS / NT;     // threads along M
    constexpr int TM = BM / MT;           // pixels per thread
    constexpr int F4_PER_PIX = BK / 4;
    constexpr int PIX_PER_PASS = NTHREADS / F4_PER_PIX;
    constexpr int A_PASSES = BM / PIX_PER_PASS;
    constexpr int LDA = BM + 4;
    constexpr int B_F4 = BK * BN / 4;

    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- per-thread A-gather coordinates (fixed over the K loop) ----
    const int a_kq = tid % F4_PER_PIX;
    int a_b[A_PASSES], a_oy[A_PASSES], a_ox[A_PASSES];
    bool a_ok[A_PASSES];
#pragma unroll
    for (int p = 0; p < A_PASSES; ++p) {
        int64_t m = m0 + p * PIX_PER_PASS + tid / F4_PER_PIX;
        a_ok[p] = m < a.M;
        int64_t mm = a_ok[p] ? m : 0;
        a_ox[p] = (int)(mm % a.Wo);
        int64_t r = mm / a.Wo;
        a_oy[p] = (int)(r % a.Ho);
        a_b[p] = (int)(r / a.Ho);
    }
    const int b_row = tid / NT, b_col4 = tid % NT;

    const int kchunks = a.Cin / BK;
    const int nk = num_taps<KIND>() * kchunks;

    float4 a_reg[A_PASSES];
    float4 b_reg = make_float4(0.f, 0.f, 0.f, 0.f);

    auto load_regs = [&](int it) {
        const int tap = it / kchunks;
        const int c0 = (it - tap * kchunks) * BK;
#pragma unroll
        for (int p = 0; p < A_PASSES; ++p) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_ok[p]) {
                int64_t pix = in_pixel<KIND>(a_b[p], a_oy[p], a_ox[p], tap, a.Hi, a.Wi);
                v = __ldg(reinterpret_cast<const float4*>(a.in + pix * a.Cin + c0) + a_kq);
                v.x = a.pre(v.x);
                v.y = a.pre(v.y);
                v.z = a.pre(v.z);
                v.w = a.pre(v.w);
            }
            a_reg[p] = v;
        }
        if (tid < B_F4) {
            b_reg = __ldg(reinterpret_cast<const float4*>(
                              a.w + ((int64_t)tap * a.Cin + c0 + b_row) * a.Cout + n0) +
                          b_col4);
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int p = 0; p < A_PASSES; ++p) {
            const int ml = p * PIX_PER_PASS + tid / F4_PER_PIX;
            As[buf][a_kq * 4 + 0][ml] = a_reg[p].x;
            As[buf][a_kq * 4 + 1][ml] = a_reg[p].y;
            As[buf][a_kq * 4 + 2][ml] = a_reg[p].z;
            As[buf][a_kq * 4 + 3][ml] = a_reg[p].w;
        }
        if (tid < B_F4) *reinterpret_cast<float4*>(&Bs[buf][b_row][b_col4 * 4]) = b_reg;
    };

    const int tn = tid % NT, tm = tid / NT;
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load_regs(0);
    store_smem(0);
    __syncthreads();

    for (int it = 0; it < nk; ++it) {
        const int cur = it & 1;
        const bool has_next = it + 1 < nk;
        if (has_next) load_regs(it + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[TM];
            if constexpr (TM % 4 == 0) {
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    float4 t = *reinterpret_cast<const float4*>(&As[cur][k][tm * TM + i]);
                    av[i] = t.x;
                    av[i + 1] = t.y;
                    av[i + 2] = t.z;
                    av[i + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < TM; ++i) av[i] = As[cur][k][tm * TM + i];
            }
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[cur][k][tn * 4]);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                acc[i][0] = fmaf(av[i], bv.x, acc[i][0]);
                acc[i][1] = fmaf(av[i], bv.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], bv.z, acc[i][2]);
                acc[i][3] = fmaf(av[i], bv.w, acc[i][3]);
            }
        }
        if (has_next) store_smem(cur ^ 1);
        __syncthreads();
    }

    // ---- epilogue:  acc * scale + bias (+ residual)  (conv_block.py:210-214) ----
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + tm * TM + i;
        if (m >= a.M) continue;
        const int64_t off = m * a.Cout + n0 + tn * 4;
        float4 v;
        v.x = acc[i][0] * a.scale + a.bias;
        v.y = acc[i][1] * a.scale + a.bias;
        v.z = acc[i][2] * a.scale + a.bias;
        v.w = acc[i][3] * a.scale + a.bias;
        if (a.res != nullptr) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(a.res + off));
            v.x += r.x;
            v.y += r.y;
            v.z += r.z;
            v.w += r.w;
        }
        *reinterpret_cast<float4*>(a.out + off) = v;
    }
}

template <int KIND, int BN, int BK>
int launch_one(const ConvArgs& a, cudaStream_t stream) {
    dim3 grid(ceil_div_u(a.M, BM), (unsigned)(a.Cout / BN));
    conv_f32_kernel<KIND, BN, BK><<<grid, NTHREADS, 0, stream>>>(a);
    return check_launch();
}

template <int KIND, int BN>
int launch_bk(const ConvArgs& a, cudaStream_t stream) {
    if (a.Cin % 16 == 0) return launch_one<KIND, BN, 16>(a, stream);
    return launch_one<KIND, BN, 8>(a, stream);
}

template <int KIND>
int launch_kind(const ConvArgs& a, cudaStream_t stream) {
    if (a.Cout % 64 == 0) return launch_bk<KIND, 64>(a, stream);
    if (a.Cout == 32) return launch_bk<KIND, 32>(a, stream);
    if (a.Cout == 16) return launch_bk<KIND, 16>(a, stream);
    if (a.Cout == 8) return launch_bk<KIND, 8>(a, stream);
    return VQAE_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// bicubic x2 upsample, align_corners=False, A=-0.75, clamped source indices
// (nn.Upsample in vq_ae/layers/conv.py:8).  For scale 2 the phase is 0.75 (even output index)
// or 0.25 (odd): taps (-9, 67, 225, -27)/256 resp. (-27, 225, 67, -9)/256 -- exact in fp32.
// out = bicubic(in) + bias  (bias: the skip path's bias1d, conv_block.py:212)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_taps(int o, int n, int idx[4], float w[4]) {
    const int fl = (o >> 1) - 1 + (o & 1);  // floor((o + 0.5) / 2 - 0.5)
    if (o & 1) {
        w[0] = -27.f / 256.f; w[1] = 225.f / 256.f; w[2] = 67.f / 256.f; w[3] = -9.f / 256.f;
    } else {
        w[0] = -9.f / 256.f; w[1] = 67.f / 256.f; w[2] = 225.f / 256.f; w[3] = -27.f / 256.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) idx[k] = min(max(fl - 1 + k, 0), n - 1);
}

__global__ void __launch_bounds__(256)
bicubic_up2_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t total4, int H,
                   int W, int C4, float bias) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int c4 = (int)(i % C4);
    int64_t r = i / C4;
    const int ox = (int)(r % (2 * W));
    r /= (2 * W);
    const int oy = (int)(r % (2 * H));
    const int b = (int)(r / (2 * H));
    int iy[4], ix[4];
    float wy[4], wx[4];
    cubic_taps(oy, H, iy, wy);
    cubic_taps(ox, W, ix, wx);
    const float4* src = reinterpret_cast<const float4*>(in) + (int64_t)b * H * W * C4 + c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
        float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
            const float4 v = __ldg(src + ((int64_t)iy[ky] * W + ix[kx]) * C4);
            row.x = fmaf(v.x, wx[kx], row.x);
            row.y = fmaf(v.y, wx[kx], row.y);
            row.z = fmaf(v.z, wx[kx], row.z);
            row.w = fmaf(v.w, wx[kx], row.w);
        }
        acc.x = fmaf(row.x, wy[ky], acc.x);
        acc.y = fmaf(row.y, wy[ky], acc.y);
        acc.z = fmaf(row.z, wy[ky], acc.z);
        acc.w = fmaf(row.w, wy[ky], acc.w);
    }
    acc.x += bias; acc.y += bias; acc.z += bias; acc.w += bias;
    reinterpret_cast<float4*>(out)[i] = acc;
}

}  // namespace

int conv_f32(int kind, const float* in, const float* w, float* out, const float* res, int64_t B,
             int Hi, int Wi, int Cin, int Cout, PreOp pre, float scale, float bias,
             cudaStream_t stream) {
    if (!in || !w || !out || B <= 0 || Hi <= 0 || Wi <= 0) return VQAE_ERR_BAD_ARG;
    if (Cin % 8 != 0 || Cout % 8 != 0) return VQAE_ERR_UNSUPPORTED;
    ConvArgs a;
    a.in = in; a.w = w; a.out = out; a.res = res;
    a.Hi = Hi; a.Wi = Wi; a.Cin = Cin; a.Cout = Cout;
    if (kind == CONV_2x2S2) {
        if ((Hi | Wi) & 1) return VQAE_ERR_UNSUPPORTED;
        a.Ho = Hi / 2; a.Wo = Wi / 2;
    } else {
        a.Ho = Hi; a.Wo = Wi;
    }
    a.M = B * a.Ho * a.Wo;
    a.pre = pre; a.scale = scale; a.bias = bias;
    switch (kind) {
        case CONV_1x1: return launch_kind<CONV_1x1>(a, stream);
        case CONV_2x2S2: return launch_kind<CONV_2x2S2>(a, stream);
        case CONV_3x3_CIRC: return launch_kind<CONV_3x3_CIRC>(a, stream);
    }
    return VQAE_ERR_BAD_ARG;
}

int bicubic_up2_f32(const float* in, float* out, int64_t B, int H, int W, int C, float bias,
                    cudaStream_t stream) {
    if (!in || !out || B <= 0) return VQAE_ERR_BAD_ARG;
    if (C % 4 != 0) return VQAE_ERR_UNSUPPORTED;
    const int64_t total4 = B * 2 * H * 2 * W * (C / 4);
    bicubic_up2_kernel<<<ceil_div_u(total4, 256), 256, 0, stream>>>(in, out, total4, H, W, C / 4,
                                                                    bias);
    return check_launch();
}

}  // namespace vqae
