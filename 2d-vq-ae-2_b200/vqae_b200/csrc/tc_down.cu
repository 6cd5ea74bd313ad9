// tcgen05 kernel for PreActFixupResBlock in mode 'down' (reference: vq_ae/layers/conv_block.py
// :196-216 with conv specs pre_activation_fixup.yaml:35-45): C_in -> C_out = 2 C_in, stride 2.
//
//   branch:  t1 = W1 . (elu(x + b1a) + b1b)              1x1, C_in -> C_out, full resolution
//            t2 = W2 (*) (elu(t1 + b2a) + b2b)           2x2 stride 2, C_out -> C_out
//            t3 = W3 . (elu(t2 + b3a) + b3b)             1x1
//   skip:    s  = Ws (*) (x + b1c) + b1d                  2x2 stride 2, C_in -> C_out
//   out = t3 * scale + b4 + s
//
// A 2x2 stride-2 conv is a 1x1 GEMM over space-to-depth: the four taps (ky,kx) read the four
// parity planes of the input.  One CTA tile = 8 x 16 output pixels (one M = 128 MMA tile) fed by
// 16 x 32 input pixels that the prologue scatters into four 128-pixel parity planes, so that
//   G1: D1[plane] = A1[plane] . W1^T                      4 MMA tiles (pointwise, any pixel order)
//   G2: D2 = sum_planes U[plane] . W2[plane]^T            K = 4 C_out
//   G3: D3 = V . (scale W3)^T + sum_planes As[plane] . Ws[plane]^T      (skip conv accumulates into
//                                                          the same TMEM tile; scale folded into W3)
//   out = D3 + (b4 + b1d)
// with every operand in the un-swizzled K-major canonical layout of tc_common.cuh.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace vqae {
namespace {

using namespace tc;

constexpr int DB_OH = 8, DB_OW = 16;          // output tile
constexpr int DB_PLANE = 129;                 // pixel pitch between parity planes (odd: no bank conflicts)
constexpr int DB_PIX = 4 * DB_PLANE + 1;      // 517: operand region pitch in pixels (odd)
constexpr uint32_t DB_LBO = DB_PIX * 16;
constexpr uint32_t DB_VLBO = DB_PLANE * 16;

template <int CI, int CO>
struct DownCfg {
    static constexpr int CIP = CI < 16 ? 16 : CI;      // input channels seen by the MMA
    static constexpr int KCI = CI / 8;                 // real 16-byte chunks per input pixel
    static constexpr int KCIP = CIP / 8;
    static constexpr int KCO = CO / 8;
    static constexpr int NW = (CO == 64) ? 8 : 4;
    static constexpr int WORKERS = NW * 32;
    static constexpr int THREADS = WORKERS + 32;       // + MMA warp
    static constexpr int NC = (CO == 64) ? 32 : CO;    // TMEM columns per epilogue unit
    static constexpr int UCH = NC / 8;
    static constexpr uint32_t W1_LBO = CO * 16, W1_BYTES = KCIP * W1_LBO;     // [CO x CIP]
    static constexpr uint32_t WO_LBO = CO * 16, WO_BYTES = KCO * WO_LBO;      // [CO x CO]
    static constexpr uint32_t OFF_A1 = 0;
    static constexpr uint32_t OFF_AS = OFF_A1 + KCIP * DB_LBO;
    static constexpr uint32_t OFF_U = OFF_AS + KCIP * DB_LBO;
    static constexpr uint32_t OFF_V = OFF_U + KCO * DB_LBO;
    static constexpr uint32_t OFF_W1 = OFF_V + KCO * DB_VLBO;
    static constexpr uint32_t OFF_W2 = OFF_W1 + W1_BYTES;                     // 4 planes
    static constexpr uint32_t OFF_W3 = OFF_W2 + 4 * WO_BYTES;
    static constexpr uint32_t OFF_WS = OFF_W3 + WO_BYTES;                     // 4 planes [CO x CIP]
    static constexpr uint32_t W_TOTAL = W1_BYTES + 4 * WO_BYTES + WO_BYTES + 4 * W1_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_W1 + W_TOTAL;
    static constexpr uint32_t SMEM = OFF_BAR + 64;
    static constexpr int TMEM_COLS = CO == 64 ? 512 : (CO == 32 ? 256 : 128);
    static constexpr int MIN_CTAS = CO == 64 ? 1 : (CO == 32 ? 2 : 3);
};

struct DownArgs {
    const void* x;                // NHWC fp32 or fp16 [B,H,W,CI]
    void* out;                    // NHWC, same element type [B,H/2,W/2,CO]
    const __nv_bfloat16* w;       // [W1 | W2 x4 | scale*W3 | Ws x4] in canonical [k-chunk][n][8]
    int n_tiles, H, W, tiles_x, tiles_per_img;
    float b1a, b1b, b2a, b2b, b3a, b3b, b1c, bsum;   // bsum = b4 + b1d
};

template <int CI, int CO, typename TIO>
__global__ void __launch_bounds__(DownCfg<CI, CO>::THREADS, DownCfg<CI, CO>::MIN_CTAS)
down_block_tc_kernel(DownArgs a) {
    using IO = StreamIO<TIO>;
    using Cfg = DownCfg<CI, CO>;
    constexpr int CIP = Cfg::CIP, KCI = Cfg::KCI, NW = Cfg::NW, NC = Cfg::NC, UCH = Cfg::UCH;
    constexpr int MMA_WARP = NW;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_mma = sbase + Cfg::OFF_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_BAR + 16);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // provably warp-uniform
    const uint32_t leader = lane == 0;

    if (tid == 0) {
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
    {
        const uint4* g = reinterpret_cast<const uint4*>(a.w);
        for (int i = tid; i < (int)(Cfg::W_TOTAL / 16); i += Cfg::THREADS)
            *reinterpret_cast<uint4*>(smem + Cfg::OFF_W1 + i * 16) = __ldg(g + i);
        for (int i = tid; i < (int)(Cfg::OFF_W1 / 16); i += Cfg::THREADS)
            *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t idesc = make_idesc_bf16(128, CO);
    const uint32_t tD2 = tmem_base + 4 * CO;           // D2 / D3 tile after the four D1 tiles

    const uint64_t dA1 = make_desc(sbase + Cfg::OFF_A1, DB_LBO, 128);
    const uint64_t dAS = make_desc(sbase + Cfg::OFF_AS, DB_LBO, 128);
    const uint64_t dU = make_desc(sbase + Cfg::OFF_U, DB_LBO, 128);
    const uint64_t dV = make_desc(sbase + Cfg::OFF_V, DB_VLBO, 128);
    const uint64_t dW1 = make_desc(sbase + Cfg::OFF_W1, Cfg::W1_LBO, 128);
    const uint64_t dW2 = make_desc(sbase + Cfg::OFF_W2, Cfg::WO_LBO, 128);
    const uint64_t dW3 = make_desc(sbase + Cfg::OFF_W3, Cfg::WO_LBO, 128);
    const uint64_t dWS = make_desc(sbase + Cfg::OFF_WS, Cfg::W1_LBO, 128);

    const int q4 = warp & 3, grp = (warp >> 2) & 1;
    const int row_in_tile = q4 * 32 + lane;
    const uint32_t t_lane = (uint32_t)(q4 * 32) << 16;
    const uint32_t t_col = (CO == 64) ? grp * 32 : 0;
    const int kc0 = (CO == 64) ? grp * 4 : 0;
    constexpr int F4 = NC / 4, SROW = NC + 4;
    float* stage = reinterpret_cast<float*>(smem + Cfg::OFF_U) + (warp < NW ? warp : 0) * 32 * SROW;
    uint32_t mma_phase = 0;
    const int Ho = a.H / 2, Wo = a.W / 2;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int img = tile / a.tiles_per_img;
        const int trem = tile - img * a.tiles_per_img;
        const int r0 = (trem / a.tiles_x) * DB_OH, c0 = (trem % a.tiles_x) * DB_OW;   // output coords
        const TIO* ximg = reinterpret_cast<const TIO*>(a.x) + (size_t)img * a.H * a.W * CI;
        TIO* oimg = reinterpret_cast<TIO*>(a.out) + (size_t)img * Ho * Wo * CO;

        // ---- P: 16 x 32 input pixels -> A1 = bf16(elu(x + b1a) + b1b), As = bf16(x + b1c),
        //      both scattered into the four parity planes ----
        if (warp < NW) {
            constexpr int ITEMS = 4 * 128 * KCI;
            constexpr int PB = 4;
            for (int base = tid; base < ITEMS; base += Cfg::WORKERS * PB) {
                float vv[PB][8];
                int dst[PB];
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const int id = base + u * Cfg::WORKERS;
                    dst[u] = -1;
                    if (id < ITEMS) {
                        const int ip = id / KCI, kc = id - ip * KCI;    // input pixel of the tile
                        const int iy = ip >> 5, ix = ip & 31;           // 16 rows x 32 cols
                        IO::load8(ximg + ((size_t)(2 * r0 + iy) * a.W + 2 * c0 + ix) * CI + kc * 8, vv[u]);
                        const int m = ((iy & 1) * 2 + (ix & 1)) * DB_PLANE + (iy >> 1) * DB_OW + (ix >> 1);
                        dst[u] = kc * (int)DB_LBO + m * 16;
                    }
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    if (dst[u] >= 0) {
                        *reinterpret_cast<uint4*>(smem + Cfg::OFF_A1 + dst[u]) = act_pack8(vv[u], a.b1a, a.b1b);
                        *reinterpret_cast<uint4*>(smem + Cfg::OFF_AS + dst[u]) = add_pack8(vv[u], a.b1c);
                    }
                }
            }
        }
        fence_proxy_async_smem();
        __syncthreads();

        // ---- G1: D1[plane] = A1[plane] . W1^T ----
        if (warp == MMA_WARP) {          // whole warp, warp-uniform operands; lane 0 issues
            {
                tc_fence_after_sync();
#pragma unroll
                for (int pl = 0; pl < 4; ++pl)
#pragma unroll
                    for (int ks = 0; ks < CIP / 16; ++ks)
                        umma_bf16(tmem_base + pl * CO,
                                  dA1 + (uint64_t)((pl * DB_PLANE * 16 + ks * 2 * DB_LBO) >> 4),
                                  dW1 + (uint64_t)((ks * 2 * Cfg::W1_LBO) >> 4), idesc, ks > 0, leader);
                umma_commit(bar_mma, leader);
            }
            __syncwarp();
        }
        // ---- E1: U[plane] = bf16(elu(D1 + b2a) + b2b) ----
        if (warp < NW) {
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            for (int pl = 0; pl < 4; ++pl) {
                float v[NC];
                tmem_ld<NC>(tmem_base + t_lane + pl * CO + t_col, v);
                tmem_ld_wait();
                const int m = pl * DB_PLANE + row_in_tile;
#pragma unroll
                for (int j = 0; j < UCH; ++j)
                    *reinterpret_cast<uint4*>(smem + Cfg::OFF_U + (kc0 + j) * DB_LBO + m * 16) =
                        act_pack8(v + 8 * j, a.b2a, a.b2b);
            }
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        fence_proxy_async_smem();
        __syncthreads();

        // ---- G2: D2 = sum over the four taps (= planes) of U[plane] . W2[plane]^T ----
        if (warp == MMA_WARP) {          // whole warp, warp-uniform operands; lane 0 issues
            {
                tc_fence_after_sync();
#pragma unroll
                for (int pl = 0; pl < 4; ++pl)
#pragma unroll
                    for (int ks = 0; ks < CO / 16; ++ks)
                        umma_bf16(tD2, dU + (uint64_t)((pl * DB_PLANE * 16 + ks * 2 * DB_LBO) >> 4),
                                  dW2 + (uint64_t)((pl * Cfg::WO_BYTES + ks * 2 * Cfg::WO_LBO) >> 4),
                                  idesc, (pl | ks) > 0, leader);
                umma_commit(bar_mma, leader);
            }
            __syncwarp();
        }
        // ---- E2: V = bf16(elu(D2 + b3a) + b3b) ----
        if (warp < NW) {
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            float v[NC];
            tmem_ld<NC>(tD2 + t_lane + t_col, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < UCH; ++j)
                *reinterpret_cast<uint4*>(smem + Cfg::OFF_V + (kc0 + j) * DB_VLBO + row_in_tile * 16) =
                    act_pack8(v + 8 * j, a.b3a, a.b3b);
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        fence_proxy_async_smem();
        __syncthreads();

        // ---- G3: D3 = V . (scale W3)^T + sum over planes of As[plane] . Ws[plane]^T ----
        if (warp == MMA_WARP) {          // whole warp, warp-uniform operands; lane 0 issues
            {
                tc_fence_after_sync();
#pragma unroll
                for (int ks = 0; ks < CO / 16; ++ks)
                    umma_bf16(tD2, dV + (uint64_t)((ks * 2 * DB_VLBO) >> 4),
                              dW3 + (uint64_t)((ks * 2 * Cfg::WO_LBO) >> 4), idesc, ks > 0, leader);
#pragma unroll
                for (int pl = 0; pl < 4; ++pl)
#pragma unroll
                    for (int ks = 0; ks < CIP / 16; ++ks)
                        umma_bf16(tD2, dAS + (uint64_t)((pl * DB_PLANE * 16 + ks * 2 * DB_LBO) >> 4),
                                  dWS + (uint64_t)((pl * Cfg::W1_BYTES + ks * 2 * Cfg::W1_LBO) >> 4),
                                  idesc, 1u, leader);
                umma_commit(bar_mma, leader);
            }
            __syncwarp();
        }
        // ---- E3: out = D3 + (b4 + b1d), transposed through shared memory for coalesced stores ----
        if (warp < NW) {
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            float v[NC];
            tmem_ld<NC>(tD2 + t_lane + t_col, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < F4; ++j)
                *reinterpret_cast<float4*>(stage + lane * SROW + 4 * j) =
                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            const int rsub = lane / F4, c4 = lane % F4;
#pragma unroll
            for (int k = 0; k < F4; ++k) {
                const int rr = rsub + k * (32 / F4);
                const int p = q4 * 32 + rr;                       // output pixel of the tile
                const size_t off =
                    ((size_t)(r0 + (p >> 4)) * Wo + c0 + (p & 15)) * CO + kc0 * 8 + c4 * 4;
                float4 d = *reinterpret_cast<const float4*>(stage + rr * SROW + 4 * c4);
                d.x += a.bsum; d.y += a.bsum; d.z += a.bsum; d.w += a.bsum;
                IO::store4(oimg + off, d);
            }
            __syncwarp();
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        __syncthreads();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int CI, int CO, typename TIO>
int launch_down(const DownArgs& a, int sm_count, cudaStream_t stream) {
    using Cfg = DownCfg<CI, CO>;
    auto kern = down_block_tc_kernel<CI, CO, TIO>;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)Cfg::SMEM));
        attr_set.cur() = true;
    }
    const int cap = sm_count * Cfg::MIN_CTAS;
    const int grid = a.n_tiles < cap ? a.n_tiles : cap;
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(a);
    return check_launch();
}

}  // namespace

size_t down_block_pack_elems(int CI) {
    const int CIP = CI < 16 ? 16 : CI, CO = 2 * CI;
    return (size_t)5 * CO * CIP + (size_t)5 * CO * CO;
}

int down_block_tc(const void* x, void* out, int io_dtype, const void* w_packed,
                  const float* scalars8, int64_t B, int H, int W, int CI, int sm_count,
                  cudaStream_t stream) {
    if (io_dtype != VQAE_DT_F32 && io_dtype != VQAE_DT_F16) return VQAE_ERR_UNSUPPORTED;
    if (!x || !out || !w_packed || !scalars8 || B <= 0) return VQAE_ERR_BAD_ARG;
    if (H % (2 * DB_OH) != 0 || W % (2 * DB_OW) != 0 || H <= 0 || W <= 0) return VQAE_ERR_UNSUPPORTED;
    DownArgs a;
    a.x = x; a.out = out; a.w = reinterpret_cast<const __nv_bfloat16*>(w_packed);
    a.H = H; a.W = W; a.tiles_x = (W / 2) / DB_OW; a.tiles_per_img = ((H / 2) / DB_OH) * a.tiles_x;
    const int64_t nt = B * a.tiles_per_img;
    if (nt > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    a.n_tiles = (int)nt;
    a.b1a = scalars8[0]; a.b1b = scalars8[1]; a.b2a = scalars8[2]; a.b2b = scalars8[3];
    a.b3a = scalars8[4]; a.b3b = scalars8[5]; a.b1c = scalars8[6]; a.bsum = scalars8[7];
    const bool h = io_dtype == VQAE_DT_F16;
    switch (CI) {
        case 32: return h ? launch_down<32, 64, __half>(a, sm_count, stream)
                          : launch_down<32, 64, float>(a, sm_count, stream);
        case 16: return h ? launch_down<16, 32, __half>(a, sm_count, stream)
                          : launch_down<16, 32, float>(a, sm_count, stream);
        case 8: return h ? launch_down<8, 16, __half>(a, sm_count, stream)
                         : launch_down<8, 16, float>(a, sm_count, stream);
    }
    return VQAE_ERR_UNSUPPORTED;
}

}  // namespace vqae
