// High-resolution half of a PreActFixupResBlock in mode 'up' (vq_ae/layers/conv_block.py:196-216 with
// ResizeConv2D, vq_ae/layers/conv.py:4-11), fused into one kernel:
//
//     out = W3 . (elu(bicubic_x2(t2) + b3a) + b3b) * scale + b4  +  (bicubic_x2(s1) + b1d)
//
// where t2 = W2 . (...) and s1 = Ws . (x + b1c) were computed at LOW resolution (a 1x1 conv without
// bias commutes with the upsample, abi.cu).  The unfused form writes the upsampled branch (4 x c_branch
// channels per low-res pixel) and the upsampled skip to HBM and reads them back; here both stay on
// chip and the only high-resolution traffic is the output.  The arithmetic is that of
// bicubic_up2_kernel + conv_f32_kernel<CONV_1x1> (same tap weights, same fmaf order over taps and
// over input channels, same epilogue expression), so the fp32 path stays bit-identical.
//
// CTA = 8 x 64 output pixels; a thread owns two horizontally adjacent pixels, which share four of their
// five source columns.  The 8 x 36 low-resolution window (clamped at the image border, like
// nn.Upsample(bicubic, align_corners=False)) is staged in shared memory with a pixel pitch of C + 4
// floats, so that the row-per-lane 128-bit reads are bank-conflict-free.
#include "common.cuh"
#include "kernels.cuh"

namespace vqae {
namespace {

constexpr int UT_TY = 8, UT_TX = 64;                      // output pixels per CTA
constexpr int UT_RY = UT_TY / 2 + 4, UT_RX = UT_TX / 2 + 4;   // low-res window (2 + 2 halo)
constexpr int UT_THREADS = UT_TY * UT_TX / 2;

struct UpTailArgs {
    const float* t2;      // [B,H,W,CB]  low-res branch (after the 1x1 of branch_conv2)
    const float* s1;      // [B,H,W,CO]  low-res skip conv output (without bias1d)
    const float* w3;      // [CB][CO]    packed branch_conv3
    float* out;           // [B,2H,2W,CO]
    int H, W;
    float b3a, b3b, scale, b4, b1d;
};

template <int CB, int CO>
struct UpTailCfg {
    static constexpr int PT = CB + 4, PS = CO + 4;        // pixel pitches in floats
    static constexpr int OFF_S = UT_RY * UT_RX * PT;      // floats
    static constexpr int OFF_W = OFF_S + UT_RY * UT_RX * PS;
    static constexpr size_t SMEM = (size_t)(OFF_W + CB * CO) * sizeof(float);
    // resident CTAs per SM the register allocation is tuned for (the kernel is latency-bound on
    // shared-memory reads: more warps in flight matter more than a few spilled values)
    static constexpr int MIN_CTAS = CB == 16 ? 4 : (CB == 32 ? 3 : 1);
};

template <int CB, int CO>
__global__ void __launch_bounds__(UT_THREADS, UpTailCfg<CB, CO>::MIN_CTAS)
up_tail_f32_kernel(UpTailArgs a) {
    using Cfg = UpTailCfg<CB, CO>;
    constexpr int PT = Cfg::PT, PS = Cfg::PS;
    extern __shared__ __align__(16) float sm[];
    float* Ts = sm;
    float* Ss = sm + Cfg::OFF_S;
    float* Ws = sm + Cfg::OFF_W;

    const int tid = threadIdx.x;
    const int oy0 = blockIdx.y * UT_TY, ox0 = blockIdx.x * UT_TX, b = blockIdx.z;
    const int ly0 = oy0 / 2 - 2, lx0 = ox0 / 2 - 2;      // low-res origin of the window

    // ---- stage the window (clamped source indices), the skip window and W3 ----
    const float* t2 = a.t2 + (size_t)b * a.H * a.W * CB;
    const float* s1 = a.s1 + (size_t)b * a.H * a.W * CO;
    for (int i = tid; i < UT_RY * UT_RX * (CB / 4); i += UT_THREADS) {
        const int c4 = i % (CB / 4), p = i / (CB / 4);
        const int ry = p / UT_RX, rx = p % UT_RX;
        const int y = min(max(ly0 + ry, 0), a.H - 1), x = min(max(lx0 + rx, 0), a.W - 1);
        *reinterpret_cast<float4*>(Ts + p * PT + 4 * c4) =
            __ldg(reinterpret_cast<const float4*>(t2 + ((size_t)y * a.W + x) * CB) + c4);
    }
    for (int i = tid; i < UT_RY * UT_RX * (CO / 4); i += UT_THREADS) {
        const int c4 = i % (CO / 4), p = i / (CO / 4);
        const int ry = p / UT_RX, rx = p % UT_RX;
        const int y = min(max(ly0 + ry, 0), a.H - 1), x = min(max(lx0 + rx, 0), a.W - 1);
        *reinterpret_cast<float4*>(Ss + p * PS + 4 * c4) =
            __ldg(reinterpret_cast<const float4*>(s1 + ((size_t)y * a.W + x) * CO) + c4);
    }
    for (int i = tid; i < CB * CO / 4; i += UT_THREADS)
        reinterpret_cast<float4*>(Ws)[i] = __ldg(reinterpret_cast<const float4*>(a.w3) + i);
    __syncthreads();

    // ---- this thread's pixel pair: (oy, ox) and (oy, ox + 1), ox even ----
    const int pr = tid / (UT_TX / 2), pc = tid % (UT_TX / 2);
    const int oy = oy0 + pr, ox = ox0 + 2 * pc;
    // cubic taps for scale 2 (conv_f32.cu, cubic_taps): even output index -> phase 0.75 on source
    // columns fl-1 .. fl+2 with fl = o/2 - 1; odd -> phase 0.25 with fl = o/2.  Window-local indices:
    const int wy0 = (oy >> 1) - 1 + (oy & 1) - 1 - ly0;   // first of the four source rows
    const int wx0 = pc;                                   // ox/2 - 2 - lx0: first of the FIVE columns
    const float we[4] = {-9.f / 256.f, 67.f / 256.f, 225.f / 256.f, -27.f / 256.f};    // even index
    const float wo[4] = {-27.f / 256.f, 225.f / 256.f, 67.f / 256.f, -9.f / 256.f};    // odd index
    float wy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wy[k] = (oy & 1) ? wo[k] : we[k];

    // bicubic value of 4 channels at both pixels: rows first (kx ascending), then ky ascending
    auto bicubic4 = [&](const float* src, int pitch, int c, float4& e, float4& o) {
        e = make_float4(0.f, 0.f, 0.f, 0.f);
        o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const float* rowp = src + ((wy0 + ky) * UT_RX + wx0) * pitch + c;
            float4 v[5];
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) v[kx] = *reinterpret_cast<const float4*>(rowp + kx * pitch);
            float4 re = make_float4(0.f, 0.f, 0.f, 0.f), ro = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                re.x = fmaf(v[kx].x, we[kx], re.x); re.y = fmaf(v[kx].y, we[kx], re.y);
                re.z = fmaf(v[kx].z, we[kx], re.z); re.w = fmaf(v[kx].w, we[kx], re.w);
                ro.x = fmaf(v[kx + 1].x, wo[kx], ro.x); ro.y = fmaf(v[kx + 1].y, wo[kx], ro.y);
                ro.z = fmaf(v[kx + 1].z, wo[kx], ro.z); ro.w = fmaf(v[kx + 1].w, wo[kx], ro.w);
            }
            e.x = fmaf(re.x, wy[ky], e.x); e.y = fmaf(re.y, wy[ky], e.y);
            e.z = fmaf(re.z, wy[ky], e.z); e.w = fmaf(re.w, wy[ky], e.w);
            o.x = fmaf(ro.x, wy[ky], o.x); o.y = fmaf(ro.y, wy[ky], o.y);
            o.z = fmaf(ro.z, wy[ky], o.z); o.w = fmaf(ro.w, wy[ky], o.w);
        }
    };

    float acc[2][CO];
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[0][j] = acc[1][j] = 0.f;
    const PreOp pre3{a.b3a, a.b3b, 1};
#pragma unroll 1
    for (int c = 0; c < CB; c += 4) {
        float4 e, o;
        bicubic4(Ts, PT, c, e, o);
        // bicubic_up2_kernel adds its bias (0 on the branch) before the conv's pre-activation
        const float ae[4] = {pre3(e.x + 0.f), pre3(e.y + 0.f), pre3(e.z + 0.f), pre3(e.w + 0.f)};
        const float ao[4] = {pre3(o.x + 0.f), pre3(o.y + 0.f), pre3(o.z + 0.f), pre3(o.w + 0.f)};
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const float* wr = Ws + (c + cc) * CO;
#pragma unroll
            for (int j = 0; j < CO; j += 4) {
                const float4 w = *reinterpret_cast<const float4*>(wr + j);
                acc[0][j] = fmaf(ae[cc], w.x, acc[0][j]); acc[0][j + 1] = fmaf(ae[cc], w.y, acc[0][j + 1]);
                acc[0][j + 2] = fmaf(ae[cc], w.z, acc[0][j + 2]); acc[0][j + 3] = fmaf(ae[cc], w.w, acc[0][j + 3]);
                acc[1][j] = fmaf(ao[cc], w.x, acc[1][j]); acc[1][j + 1] = fmaf(ao[cc], w.y, acc[1][j + 1]);
                acc[1][j + 2] = fmaf(ao[cc], w.z, acc[1][j + 2]); acc[1][j + 3] = fmaf(ao[cc], w.w, acc[1][j + 3]);
            }
        }
    }

    // ---- epilogue: acc * scale + bias4 + (bicubic(skip) + bias1d), 128-bit stores ----
    float* outp = a.out + (((size_t)b * 2 * a.H + oy) * 2 * a.W + ox) * CO;
#pragma unroll
    for (int j = 0; j < CO; j += 4) {
        float4 se, so;
        bicubic4(Ss, PS, j, se, so);
        float4 ve, vo;
        ve.x = acc[0][j] * a.scale + a.b4;     ve.y = acc[0][j + 1] * a.scale + a.b4;
        ve.z = acc[0][j + 2] * a.scale + a.b4; ve.w = acc[0][j + 3] * a.scale + a.b4;
        vo.x = acc[1][j] * a.scale + a.b4;     vo.y = acc[1][j + 1] * a.scale + a.b4;
        vo.z = acc[1][j + 2] * a.scale + a.b4; vo.w = acc[1][j + 3] * a.scale + a.b4;
        ve.x += se.x + a.b1d; ve.y += se.y + a.b1d; ve.z += se.z + a.b1d; ve.w += se.w + a.b1d;
        vo.x += so.x + a.b1d; vo.y += so.y + a.b1d; vo.z += so.z + a.b1d; vo.w += so.w + a.b1d;
        *reinterpret_cast<float4*>(outp + j) = ve;
        *reinterpret_cast<float4*>(outp + CO + j) = vo;
    }
}

template <int CB, int CO>
int launch_up_tail(const UpTailArgs& a, int64_t B, cudaStream_t stream) {
    using Cfg = UpTailCfg<CB, CO>;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(up_tail_f32_kernel<CB, CO>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set.cur() = true;
    }
    dim3 grid(2 * a.W / UT_TX, 2 * a.H / UT_TY, (unsigned)B);
    up_tail_f32_kernel<CB, CO><<<grid, UT_THREADS, Cfg::SMEM, stream>>>(a);
    return check_launch();
}

}  // namespace

bool up_tail_supported(int64_t B, int H, int W, int cb, int co) {
    return B > 0 && B <= 65535 && co * 2 == cb && (cb == 16 || cb == 32 || cb == 64) &&
           (2 * W) % UT_TX == 0 && (2 * H) % UT_TY == 0;
}

int up_tail_f32(const float* t2, const float* s1, const float* w3, float* out, int64_t B, int H, int W,
                int cb, int co, float b3a, float b3b, float scale, float b4, float b1d,
                cudaStream_t stream) {
    if (!t2 || !s1 || !w3 || !out) return VQAE_ERR_BAD_ARG;
    if (!up_tail_supported(B, H, W, cb, co)) return VQAE_ERR_UNSUPPORTED;
    UpTailArgs a{t2, s1, w3, out, H, W, b3a, b3b, scale, b4, b1d};
    switch (cb) {
        case 16: return launch_up_tail<16, 8>(a, B, stream);
        case 32: return launch_up_tail<32, 16>(a, B, stream);
        case 64: return launch_up_tail<64, 32>(a, B, stream);
    }
    return VQAE_ERR_UNSUPPORTED;
}

}  // namespace vqae
