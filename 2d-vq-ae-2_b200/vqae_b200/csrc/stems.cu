// Input normalisation and the two stem convolutions (HBM-bound, CUDA cores).
//   normalize_u8   : albumentations Normalize + ToTensorV2
//                    (conf/transforms/camelyon16_transforms.yaml:1-23, transforms/normalize.yaml)
//   stem_in_f32    : Encoder.in_stem  (vq_ae/model.py:141,198)  3x3, zero pad, bias, 3 -> 8
//   stem_out_f32   : Decoder.out_stem (vq_ae/model.py:291)      3x3, zero pad, bias, 8 -> 3
#include "common.cuh"
#include "kernels.cuh"

namespace vqae {
namespace {

struct Norm3 {
    float sub[3];  // 255 * mean_c
    float mul[3];  // 1 / (255 * std_c)
};

// same operation order as the reference's numpy code: (x - mean255) * (1 / std255), no FMA
__device__ __forceinline__ float norm_px(uint8_t v, float sub, float mul) {
    return __fmul_rn(__fsub_rn((float)v, sub), mul);
}

inline Norm3 make_norm(const float* mean, const float* stdv) {
    Norm3 n;
    for (int c = 0; c < 3; ++c) {
        n.sub[c] = mean[c] * 255.0f;
        n.mul[c] = 1.0f / (stdv[c] * 255.0f);
    }
    return n;
}

__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int64_t npix,
                    int64_t hw, Norm3 n, int out_layout) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const uint8_t* s = img + p * 3;
    const float r = norm_px(s[0], n.sub[0], n.mul[0]);
    const float g = norm_px(s[1], n.sub[1], n.mul[1]);
    const float b = norm_px(s[2], n.sub[2], n.mul[2]);
    if (out_layout == VQAE_LAYOUT_NHWC) {
        out[p * 3 + 0] = r;
        out[p * 3 + 1] = g;
        out[p * 3 + 2] = b;
    } else {
        const int64_t bi = p / hw, q = p % hw;
        float* o = out + bi * 3 * hw + q;
        o[0] = r;
        o[hw] = g;
        o[2 * hw] = b;
    }
}

// XKIND: 0 = fp32 NCHW, 1 = fp32 NHWC, 2 = u8 NHWC (normalised on the fly)
template <int XKIND>
__global__ void __launch_bounds__(256)
stem_in_kernel(const void* __restrict__ xv, const float* __restrict__ w_oihw,
               const float* __restrict__ bias, float* __restrict__ out, int64_t npix, int H,
               int W, Norm3 n) {
    __shared__ float ws[27][8];  // [(c*9 + ky*3 + kx)][o]
    __shared__ float bs[8];
    for (int i = threadIdx.x; i < 27 * 8; i += blockDim.x) {
        const int o = i / 27, r = i % 27;  // OIHW: o*27 + c*9 + ky*3 + kx
        ws[r][o] = w_oihw[i];
    }
    if (threadIdx.x < 8) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();

    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int x = (int)(p % W);
    const int64_t r = p / W;
    const int y = (int)(r % H);
    const int64_t b = r / H;
    const int64_t hw = (int64_t)H * W;

    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = bs[o];

#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = x + kx - 1;
            if (ix < 0 || ix >= W) continue;
            float v[3];
            if (XKIND == 0) {
                const float* s = reinterpret_cast<const float*>(xv) + b * 3 * hw +
                                 (int64_t)iy * W + ix;
                v[0] = __ldg(s);
                v[1] = __ldg(s + hw);
                v[2] = __ldg(s + 2 * hw);
            } else if (XKIND == 1) {
                const float* s =
                    reinterpret_cast<const float*>(xv) + (b * hw + (int64_t)iy * W + ix) * 3;
                v[0] = __ldg(s);
                v[1] = __ldg(s + 1);
                v[2] = __ldg(s + 2);
            } else {
                const uint8_t* s =
                    reinterpret_cast<const uint8_t*>(xv) + (b * hw + (int64_t)iy * W + ix) * 3;
                v[0] = norm_px(__ldg(s), n.sub[0], n.mul[0]);
                v[1] = norm_px(__ldg(s + 1), n.sub[1], n.mul[1]);
                v[2] = norm_px(__ldg(s + 2), n.sub[2], n.mul[2]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* wr = ws[c * 9 + ky * 3 + kx];
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[o] = fmaf(v[c], wr[o], acc[o]);
            }
        }
    }
    float4* o4 = reinterpret_cast<float4*>(out + p * 8);
    o4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

template <int CIN>
__global__ void __launch_bounds__(256)
stem_out_kernel(const float* __restrict__ x, const float* __restrict__ w_oihw,
                const float* __restrict__ bias, float* __restrict__ out, int out_layout,
                int64_t npix, int H, int W) {
    __shared__ float ws[9 * CIN][3];  // [(ky*3+kx)*CIN + c][o]
    __shared__ float bs[3];
    for (int i = threadIdx.x; i < 3 * CIN * 9; i += blockDim.x) {
        const int o = i / (CIN * 9);
        const int c = (i / 9) % CIN;
        const int t = i % 9;
        ws[t * CIN + c][o] = w_oihw[i];
    }
    if (threadIdx.x < 3) bs[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();

    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int xx = (int)(p % W);
    const int64_t r = p / W;
    const int y = (int)(r % H);
    const int64_t b = r / H;
    const int64_t hw = (int64_t)H * W;

    float a0 = bs[0], a1 = bs[1], a2 = bs[2];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = xx + kx - 1;
            if (ix < 0 || ix >= W) continue;
            const float4* s =
                reinterpret_cast<const float4*>(x + (b * hw + (int64_t)iy * W + ix) * CIN);
#pragma unroll
            for (int c4 = 0; c4 < CIN / 4; ++c4) {
                const float4 v = __ldg(s + c4);
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float* wr = ws[(ky * 3 + kx) * CIN + c4 * 4 + j];
                    a0 = fmaf(vv[j], wr[0], a0);
                    a1 = fmaf(vv[j], wr[1], a1);
                    a2 = fmaf(vv[j], wr[2], a2);
                }
            }
        }
    }
    if (out_layout == VQAE_LAYOUT_NHWC) {
        out[p * 3 + 0] = a0;
        out[p * 3 + 1] = a1;
        out[p * 3 + 2] = a2;
    } else {
        float* o = out + b * 3 * hw + (int64_t)y * W + xx;
        o[0] = a0;
        o[hw] = a1;
        o[2 * hw] = a2;
    }
}

}  // namespace

int normalize_u8(const uint8_t* img, float* out, int64_t B, int H, int W, const float* mean,
                 const float* stdv, int out_layout, cudaStream_t stream) {
    if (!img || !out || !mean || !stdv || B <= 0 || H <= 0 || W <= 0) return VQAE_ERR_BAD_ARG;
    const int64_t hw = (int64_t)H * W, npix = B * hw;
    normalize_u8_kernel<<<ceil_div_u(npix, 256), 256, 0, stream>>>(img, out, npix, hw,
                                                                   make_norm(mean, stdv),
                                                                   out_layout);
    return check_launch();
}

int stem_in_f32(const void* x, int x_dtype, int x_layout, const float* w, const float* bias,
                float* out, int64_t B, int H, int W, int c_out, const float* mean,
                const float* stdv, cudaStream_t stream) {
    if (!x || !w || !bias || !out || B <= 0 || H <= 0 || W <= 0) return VQAE_ERR_BAD_ARG;
    if (c_out != 8) return VQAE_ERR_UNSUPPORTED;
    const int64_t npix = B * H * W;
    const unsigned grid = ceil_div_u(npix, 256);
    Norm3 n{};
    if (x_dtype == VQAE_DT_U8) {
        if (!mean || !stdv) return VQAE_ERR_BAD_ARG;
        if (x_layout != VQAE_LAYOUT_NHWC) return VQAE_ERR_UNSUPPORTED;
        n = make_norm(mean, stdv);
        stem_in_kernel<2><<<grid, 256, 0, stream>>>(x, w, bias, out, npix, H, W, n);
    } else if (x_dtype == VQAE_DT_F32) {
        if (x_layout == VQAE_LAYOUT_NCHW)
            stem_in_kernel<0><<<grid, 256, 0, stream>>>(x, w, bias, out, npix, H, W, n);
        else
            stem_in_kernel<1><<<grid, 256, 0, stream>>>(x, w, bias, out, npix, H, W, n);
    } else {
        return VQAE_ERR_UNSUPPORTED;
    }
    return check_launch();
}

int stem_out_f32(const float* x, const float* w, const float* bias, float* out, int out_layout,
                 int64_t B, int H, int W, int c_in, cudaStream_t stream) {
    if (!x || !w || !bias || !out || B <= 0 || H <= 0 || W <= 0) return VQAE_ERR_BAD_ARG;
    if (c_in != 8) return VQAE_ERR_UNSUPPORTED;
    const int64_t npix = B * H * W;
    stem_out_kernel<8><<<ceil_div_u(npix, 256), 256, 0, stream>>>(x, w, bias, out, out_layout,
                                                                  npix, H, W);
    return check_launch();
}

}  // namespace vqae
