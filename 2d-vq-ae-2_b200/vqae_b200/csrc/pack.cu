// Weight packing: OIHW fp32 conv weights of the reference's modules -> the layouts the kernels read.
// One kernel serves every format; vqae_pack_batched packs ALL convs of a model (a table of
// descriptors in device memory, one grid row per descriptor) in one launch, the single-matrix
// entry points of the ABI pass one descriptor by value.
//
// Formats (include/vqae_b200.h, VQAE_PACK_*):
//   F32_CONV      [tap][C_in][C_out] fp32                                  (conv_f32.cu)
//   SAME_BF16     11 matrices [W1 | W2 tap 0..8 | W3], each UMMA canonical K-major [k-chunk][n][8],
//                 zero padded from c to max(c, 16) channels               (tc_kernels.cu, tc_chain.cu)
//   RESIDENT_BF16 the same with branch_conv3 pre-multiplied by the Fixup scale      (tc_resident.cu)
//   DOWN_BF16     [W1 | W2 x4 planes | scale*W3 | Ws x4 planes], C_in zero padded to >= 16 (tc_down.cu)
// kind | VQAE_PACK_LO holds f16(w - f16(w)): the low half of the split operands of the fp32-accurate
// tensor-core mode (tc_split.cu, mma_down.cu SPLIT); there the matrices are first multiplied by the
// powers of two vqae_pack_desc.premul[] (branch_conv1, 2, 3, skip; 0 = 1) so that the low halves stay
// normal fp16 numbers.
#include "common.cuh"
#include "kernels.cuh"

#include "tc_common.cuh"

namespace vqae {
namespace {

// fp32 -> operand element (tc_common.cuh: fp16, or bf16 when VQAE_OPERAND_F16 == 0); lo = the
// rounding remainder, again rounded to the operand type
#if VQAE_OPERAND_F16
using op_t = __half;
__device__ __forceinline__ op_t op_round(float v) { return __float2half_rn(v); }
__device__ __forceinline__ float op_float(op_t v) { return __half2float(v); }
#else
using op_t = __nv_bfloat16;
__device__ __forceinline__ op_t op_round(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float op_float(op_t v) { return __bfloat162float(v); }
#endif
__device__ __forceinline__ op_t to_bf16(float v, bool lo) {
    const op_t hi = op_round(v);
    return lo ? op_round(v - op_float(hi)) : hi;
}

// canonical [k-chunk][n][8] index r of an N-row matrix -> (n, k)
__device__ __forceinline__ void canon(int r, int N, int& n, int& k) {
    const int kc = r / (N * 8);
    n = (r / 8) % N;
    k = kc * 8 + (r % 8);
}

__device__ void pack_element(const vqae_pack_desc& d, int i) {
    const float* w1 = reinterpret_cast<const float*>(d.src[0]);
    const float* w2 = reinterpret_cast<const float*>(d.src[1]);
    const float* w3 = reinterpret_cast<const float*>(d.src[2]);
    const float* ws = reinterpret_cast<const float*>(d.src[3]);
    const int kind = d.kind & 0xff;
    const bool lo = (d.kind & VQAE_PACK_LO) != 0;
    const float pm1 = d.premul[0] > 0.f ? d.premul[0] : 1.f, pm2 = d.premul[1] > 0.f ? d.premul[1] : 1.f;
    const float pm3 = d.premul[2] > 0.f ? d.premul[2] : 1.f, pms = d.premul[3] > 0.f ? d.premul[3] : 1.f;
    if (kind == VQAE_PACK_F32_CONV) {
        // packed index i = (t * I + c) * O + o
        const int O = d.c_out, I = d.c_in, taps = d.taps;
        const int o = i % O, c = (i / O) % I, t = i / (O * I);
        reinterpret_cast<float*>(d.dst)[i] = w1[((int64_t)o * I + c) * taps + t];
        return;
    }
    op_t* out = reinterpret_cast<op_t*>(d.dst);
    if (kind == VQAE_PACK_SAME_F16 || kind == VQAE_PACK_RESIDENT_F16) {
        const int CR = d.c_in, CP = CR < 16 ? 16 : CR;
        const int per = CP * CP;
        const int m = i / per;
        int n, k;
        canon(i % per, CP, n, k);
        float v = 0.f;
        if (n < CR && k < CR) {
            if (m == 0) v = w1[n * CR + k] * pm1;
            else if (m == 10) v = w3[n * CR + k] * (kind == VQAE_PACK_RESIDENT_F16 ? d.scale : 1.f) * pm3;
            else v = w2[((size_t)n * CR + k) * 9 + (m - 1)] * pm2;
        }
        out[i] = to_bf16(v, lo);
        return;
    }
    if (kind == VQAE_PACK_SAME_MMA_F16) {
        // [11][n][k] row-major: W1 | W2 tap 0..8 | W3 (mma_same.cu reads B fragments from rows n)
        // For c >= 16 the kernel reads x and writes out with one 128-bit access per lane: lane
        // (g, t) holds channels 4t .. 4t + 3 of a 16-channel group in fragment slots 2t, 2t + 1,
        // 2t + 8, 2t + 9.  Slot i of a group therefore stands for channel perm(i); W1's input (k)
        // order and W3's output (n) order are stored in slot order.
        const int Cc = d.c_in, per = Cc * Cc;
        const int m = i / per, n = (i % per) / Cc, k = i % Cc;
        auto perm = [Cc](int s) {
            if (Cc < 16) return s;
            const int r = s & 15;
            return (s & ~15) + 4 * ((r & 7) >> 1) + 2 * (r >> 3) + (r & 1);
        };
        float v;
        if (m == 0) v = w1[n * Cc + perm(k)] * pm1;
        else if (m == 10) v = w3[perm(n) * Cc + k] * pm3;
        else v = w2[((size_t)n * Cc + k) * 9 + (m - 1)] * pm2;
        out[i] = to_bf16(v, lo);
        return;
    }
    if (kind == VQAE_PACK_DOWN_MMA_F16) {
        // dense [n][k] rows: W1 [CO][CI] | W2 x4 [CO][CO] | scale*W3 [CO][CO] | Ws x4 [CO][CI]
        // (mma_down.cu).  Slot order as in SAME_MMA: input channels of W1 / Ws (c_in >= 16) and
        // output channels of W3 / Ws are stored in fragment-slot order.
        const int CI = d.c_in, CO = d.c_out;
        auto perm = [](int s, int width) {
            if (width < 16) return s;
            const int r = s & 15;
            return (s & ~15) + 4 * ((r & 7) >> 1) + 2 * (r >> 3) + (r & 1);
        };
        const int n1 = CO * CI, n2 = 4 * CO * CO, n3 = CO * CO;
        int j = i;
        float v;
        if (j < n1) {
            v = w1[(j / CI) * CI + perm(j % CI, CI)] * pm1;
        } else if ((j -= n1) < n2) {
            const int tap = j / (CO * CO), r = j % (CO * CO);
            v = w2[((size_t)(r / CO) * CO + r % CO) * 4 + tap] * pm2;
        } else if ((j -= n2) < n3) {
            v = w3[perm(j / CO, CO) * CO + j % CO] * d.scale * pm3;
        } else {
            j -= n3;
            const int tap = j / n1, r = j % n1;
            v = ws[((size_t)perm(r / CI, CO) * CI + perm(r % CI, CI)) * 4 + tap] * pms;
        }
        out[i] = to_bf16(v, lo);
        return;
    }
    if (kind == VQAE_PACK_UP_MMA_F16) {
        // dense [n][k] rows: W1 [CB][CI] | W2 [CB][CB] | Ws [CO][CI] | scale*W3 [CO][CB], CB = CI,
        // CO = CI / 2 (mma_up.cu).  Slot order: k of W1 / Ws / W3 (their A fragments come from
        // 128-bit channel groups), n of W2 / Ws / W3 (128-bit stores) where the width is >= 16.
        const int CI = d.c_in, CB = d.c_in, CO = d.c_out;
        auto perm = [](int s, int width) {
            if (width < 16) return s;
            const int r = s & 15;
            return (s & ~15) + 4 * ((r & 7) >> 1) + 2 * (r >> 3) + (r & 1);
        };
        const int n1 = CB * CI, n2 = CB * CB, ns = CO * CI;
        int j = i;
        float v;
        if (j < n1) v = w1[(j / CI) * CI + perm(j % CI, CI)];
        else if ((j -= n1) < n2) v = w2[perm(j / CB, CB) * CB + j % CB];
        else if ((j -= n2) < ns) v = ws[perm(j / CI, CO) * CI + perm(j % CI, CI)];
        else { j -= ns; v = w3[perm(j / CB, CO) * CB + perm(j % CB, CB)] * d.scale; }
        out[i] = to_bf16(v, lo);
        return;
    }
    if (kind == VQAE_PACK_DOWN_F16) {
        const int CI = d.c_in, CIP = CI < 16 ? 16 : CI, CO = d.c_out;
        const int n1 = CO * CIP, no = CO * CO;
        int n, k, j = i;
        float v = 0.f;
        if (j < n1) {                                         // W1 [CO x CIP]  (OIHW 1x1)
            canon(j, CO, n, k);
            if (k < CI) v = w1[n * CI + k];
        } else if ((j -= n1) < 4 * no) {                      // W2 planes: w2[n][k][ky][kx]
            const int pl = j / no;
            canon(j % no, CO, n, k);
            v = w2[((size_t)n * CO + k) * 4 + pl];
        } else if ((j -= 4 * no) < no) {                      // scale * W3
            canon(j, CO, n, k);
            v = w3[n * CO + k] * d.scale;
        } else {                                              // Ws planes: ws[n][k][ky][kx], k < CI
            j -= no;
            const int pl = j / n1;
            canon(j % n1, CO, n, k);
            if (k < CI) v = ws[((size_t)n * CI + k) * 4 + pl];
        }
        out[i] = to_bf16(v, lo);
    }
}

__global__ void __launch_bounds__(256) pack_batched_kernel(const vqae_pack_desc* __restrict__ descs) {
    const vqae_pack_desc d = descs[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n_elems; i += gridDim.x * blockDim.x)
        pack_element(d, i);
}

__global__ void __launch_bounds__(256) pack_one_kernel(const vqae_pack_desc d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d.n_elems) pack_element(d, i);
}

}  // namespace

size_t pack_elems(int kind, int c_in, int c_out, int taps) {
    switch (kind & 0xff) {
        case VQAE_PACK_F32_CONV: return (size_t)c_in * c_out * taps;
        case VQAE_PACK_SAME_F16:
        case VQAE_PACK_RESIDENT_F16: {
            const size_t cp = c_in < 16 ? 16 : c_in;
            return 11 * cp * cp;
        }
        case VQAE_PACK_SAME_MMA_F16: return (size_t)11 * c_in * c_in;
        case VQAE_PACK_UP_MMA_F16:
            return 2 * (size_t)c_in * c_in + 2 * (size_t)c_out * c_in;
        case VQAE_PACK_DOWN_MMA_F16:
            return (size_t)c_out * c_in + 4 * (size_t)c_out * c_out + (size_t)c_out * c_out +
                   4 * (size_t)c_out * c_in;
        case VQAE_PACK_DOWN_F16: {
            const size_t cip = c_in < 16 ? 16 : c_in, co = c_out;
            return co * cip + 4 * co * co + co * co + 4 * co * cip;
        }
    }
    return 0;
}

int pack_one(vqae_pack_desc d, cudaStream_t stream) {
    if (!d.src[0] || !d.dst) return VQAE_ERR_BAD_ARG;
    const size_t n = pack_elems(d.kind, d.c_in, d.c_out, d.taps);
    if (n == 0 || n > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    d.n_elems = (int)n;
    pack_one_kernel<<<ceil_div_u((int64_t)n, 256), 256, 0, stream>>>(d);
    return check_launch();
}

int pack_batched(const vqae_pack_desc* descs_dev, int n_descs, int max_elems, cudaStream_t stream) {
    if (!descs_dev || n_descs <= 0 || max_elems <= 0) return VQAE_ERR_BAD_ARG;
    if (n_descs > 65535) return VQAE_ERR_UNSUPPORTED;
    unsigned gx = ceil_div_u(max_elems, 256 * 4);
    if (gx > 64) gx = 64;
    pack_batched_kernel<<<dim3(gx, (unsigned)n_descs), 256, 0, stream>>>(descs_dev);
    return check_launch();
}

// ---- single-matrix entry points (ABI v2 names) -------------------------------------------------
static vqae_pack_desc make_desc(int kind, const float* a, const float* b, const float* c,
                                const float* d, void* dst, int c_in, int c_out, int taps,
                                float scale) {
    vqae_pack_desc p{};
    p.kind = kind; p.c_in = c_in; p.c_out = c_out; p.taps = taps; p.scale = scale;
    p.src[0] = a; p.src[1] = b; p.src[2] = c; p.src[3] = d;
    p.dst = dst;
    return p;
}

int pack_conv_weight_f32(const float* w, float* packed, int O, int I, int taps,
                         cudaStream_t stream) {
    if (!w || !packed || O <= 0 || I <= 0 || taps <= 0) return VQAE_ERR_BAD_ARG;
    return pack_one(make_desc(VQAE_PACK_F32_CONV, w, nullptr, nullptr, nullptr, packed, I, O, taps,
                              1.f), stream);
}

int pack_same_block_f16(const float* w1, const float* w2, const float* w3, int C, void* packed,
                         cudaStream_t stream) {
    if (!w1 || !w2 || !w3 || !packed) return VQAE_ERR_BAD_ARG;
    if (C != 8 && C != 16 && C != 32 && C != 64 && C != 128) return VQAE_ERR_UNSUPPORTED;
    return pack_one(make_desc(VQAE_PACK_SAME_F16, w1, w2, w3, nullptr, packed, C, C, 9, 1.f), stream);
}

int pack_resident_block_f16(const float* w1, const float* w2, const float* w3, int C, float scale,
                             void* packed, cudaStream_t stream) {
    if (!w1 || !w2 || !w3 || !packed) return VQAE_ERR_BAD_ARG;
    if (C != 32 && C != 64 && C != 128) return VQAE_ERR_UNSUPPORTED;
    return pack_one(make_desc(VQAE_PACK_RESIDENT_F16, w1, w2, w3, nullptr, packed, C, C, 9, scale),
                    stream);
}

int pack_down_block_f16(const float* w1, const float* w2, const float* w3, const float* ws, int CI,
                         float scale, void* packed, cudaStream_t stream) {
    if (!w1 || !w2 || !w3 || !ws || !packed) return VQAE_ERR_BAD_ARG;
    if (CI != 8 && CI != 16 && CI != 32 && CI != 64) return VQAE_ERR_UNSUPPORTED;
    return pack_one(make_desc(VQAE_PACK_DOWN_F16, w1, w2, w3, ws, packed, CI, 2 * CI, 4, scale),
                    stream);
}

}  // namespace vqae
