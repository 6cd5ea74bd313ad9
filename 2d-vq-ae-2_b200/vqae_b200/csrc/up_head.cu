// Low-resolution half of a PreActFixupResBlock in mode 'up' (vq_ae/layers/conv_block.py:196-216) in
// one pointwise kernel:
//     t2 = W2 . (elu(W1 . (elu(x + b1a) + b1b) + b2a) + b2b)      branch_conv1, then the 1x1 of the
//                                                                  ResizeConv2D branch_conv2 (it commutes
//                                                                  with the bicubic upsample, abi.cu)
//     s1 = Ws . (x + b1c)                                          the 1x1 of the skip ResizeConv2D
// One thread owns one low-resolution pixel: x (CB floats) and t1 stay in registers, the three weight
// matrices sit in shared memory and are read as warp-wide broadcasts.  Same arithmetic as three calls of
// conv_f32_kernel<CONV_1x1> (fmaf chain over the input channels in ascending order from 0, epilogue
// acc * 1 + 0), so the fp32 path stays bit-identical; the intermediate t1 never reaches HBM and two
// launches disappear.  up_tail.cu consumes t2 and s1.
#include "common.cuh"
#include "kernels.cuh"

namespace vqae {
namespace {

constexpr int UH_THREADS = 128;

struct UpHeadArgs {
    const float* x;       // [P][CB]
    const float* w1;      // [CB][CB] packed (input channel major)
    const float* w2;      // [CB][CB]
    const float* ws;      // [CB][CO]
    float* t2;            // [P][CB]
    float* s1;            // [P][CO]
    int64_t P;
    float b1a, b1b, b2a, b2b, b1c, one, zero;
};

template <int CB, int CO>
__global__ void __launch_bounds__(UH_THREADS)
up_head_f32_kernel(UpHeadArgs a) {
    extern __shared__ __align__(16) float sm[];
    float* W1 = sm;
    float* W2 = sm + CB * CB;
    float* WS = sm + 2 * CB * CB;
    for (int i = threadIdx.x; i < CB * CB / 4; i += UH_THREADS) {
        reinterpret_cast<float4*>(W1)[i] = __ldg(reinterpret_cast<const float4*>(a.w1) + i);
        reinterpret_cast<float4*>(W2)[i] = __ldg(reinterpret_cast<const float4*>(a.w2) + i);
    }
    for (int i = threadIdx.x; i < CB * CO / 4; i += UH_THREADS)
        reinterpret_cast<float4*>(WS)[i] = __ldg(reinterpret_cast<const float4*>(a.ws) + i);
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * UH_THREADS + threadIdx.x;
    if (p >= a.P) return;

    float x[CB];
    {
        const float4* src = reinterpret_cast<const float4*>(a.x + p * CB);
#pragma unroll
        for (int i = 0; i < CB / 4; ++i) {
            const float4 v = __ldg(src + i);
            x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
    }
    // out[o] = sum_c in(c) * W[c][o], 16 outputs at a time (accumulators in registers)
    auto gemv16 = [&](const float* W, int ldw, int o0, auto in, float (&acc)[16]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            const float v = in(c);
            const float4* wr = reinterpret_cast<const float4*>(W + c * ldw + o0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w = wr[j];
                acc[4 * j] = fmaf(v, w.x, acc[4 * j]);
                acc[4 * j + 1] = fmaf(v, w.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(v, w.z, acc[4 * j + 2]);
                acc[4 * j + 3] = fmaf(v, w.w, acc[4 * j + 3]);
            }
        }
    };
    auto store16 = [&](float* dst, const float (&acc)[16]) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            reinterpret_cast<float4*>(dst)[j] =
                make_float4(acc[4 * j] * a.one + a.zero, acc[4 * j + 1] * a.one + a.zero,
                            acc[4 * j + 2] * a.one + a.zero, acc[4 * j + 3] * a.one + a.zero);
    };
    const PreOp pre1{a.b1a, a.b1b, 1}, pre2{a.b2a, a.b2b, 1}, pre_skip{a.b1c, 0.f, 0};
    float acc[16];
    // skip: s1 = Ws . (x + b1c)
#pragma unroll
    for (int o0 = 0; o0 < CO; o0 += 16) {
        if constexpr (CO >= 16) {
            gemv16(WS, CO, o0, [&](int c) { return pre_skip(x[c]); }, acc);
            store16(a.s1 + p * CO + o0, acc);
        }
    }
    if constexpr (CO == 8) {
        float a8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a8[j] = 0.f;
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            const float v = pre_skip(x[c]);
            const float4 w0 = *reinterpret_cast<const float4*>(WS + c * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(WS + c * 8 + 4);
            a8[0] = fmaf(v, w0.x, a8[0]); a8[1] = fmaf(v, w0.y, a8[1]);
            a8[2] = fmaf(v, w0.z, a8[2]); a8[3] = fmaf(v, w0.w, a8[3]);
            a8[4] = fmaf(v, w1.x, a8[4]); a8[5] = fmaf(v, w1.y, a8[5]);
            a8[6] = fmaf(v, w1.z, a8[6]); a8[7] = fmaf(v, w1.w, a8[7]);
        }
        float4* d = reinterpret_cast<float4*>(a.s1 + p * 8);
        d[0] = make_float4(a8[0] * a.one + a.zero, a8[1] * a.one + a.zero, a8[2] * a.one + a.zero,
                           a8[3] * a.one + a.zero);
        d[1] = make_float4(a8[4] * a.one + a.zero, a8[5] * a.one + a.zero, a8[6] * a.one + a.zero,
                           a8[7] * a.one + a.zero);
    }
    // branch_conv1: t1 = W1 . pre1(x); the activation is applied once per input channel
#pragma unroll
    for (int c = 0; c < CB; ++c) x[c] = pre1(x[c]);
    float t1[CB];
#pragma unroll
    for (int o0 = 0; o0 < CB; o0 += 16) {
        gemv16(W1, CB, o0, [&](int c) { return x[c]; }, acc);
#pragma unroll
        for (int j = 0; j < 16; ++j) t1[o0 + j] = acc[j] * a.one + a.zero;   // conv epilogue: scale 1, bias 0
    }
    // 1x1 of branch_conv2 at low resolution: t2 = W2 . pre2(t1)
#pragma unroll
    for (int c = 0; c < CB; ++c) t1[c] = pre2(t1[c]);
#pragma unroll
    for (int o0 = 0; o0 < CB; o0 += 16) {
        gemv16(W2, CB, o0, [&](int c) { return t1[c]; }, acc);
        store16(a.t2 + p * CB + o0, acc);
    }
}

template <int CB, int CO>
int launch_up_head(const UpHeadArgs& a, cudaStream_t stream) {
    constexpr size_t smem = (size_t)(2 * CB * CB + CB * CO) * sizeof(float);
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(up_head_f32_kernel<CB, CO>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set.cur() = true;
    }
    up_head_f32_kernel<CB, CO><<<ceil_div_u(a.P, UH_THREADS), UH_THREADS, smem, stream>>>(a);
    return check_launch();
}

}  // namespace

bool up_head_supported(int64_t P, int ci, int cb, int co) {
    // cb = 64 instantiates (216 registers per thread) but loses to three conv_f32 launches at the
    // 32 x 32 resolution it occurs at (0.96 vs 0.74 ms for the block at batch 256): not dispatched
    return P > 0 && P <= 0x7fffffffll * UH_THREADS / 2 && ci == cb && co * 2 == cb &&
           (cb == 16 || cb == 32);
}

int up_head_f32(const float* x, const float* w1, const float* w2, const float* ws, float* t2, float* s1,
                int64_t P, int ci, int cb, int co, float b1a, float b1b, float b2a, float b2b, float b1c,
                cudaStream_t stream) {
    if (!x || !w1 || !w2 || !ws || !t2 || !s1) return VQAE_ERR_BAD_ARG;
    if (!up_head_supported(P, ci, cb, co)) return VQAE_ERR_UNSUPPORTED;
    UpHeadArgs a{x, w1, w2, ws, t2, s1, P, b1a, b1b, b2a, b2b, b1c, 1.f, 0.f};
    switch (cb) {
        case 16: return launch_up_head<16, 8>(a, stream);
        case 32: return launch_up_head<32, 16>(a, stream);
    }
    return VQAE_ERR_UNSUPPORTED;
}

}  // namespace vqae
