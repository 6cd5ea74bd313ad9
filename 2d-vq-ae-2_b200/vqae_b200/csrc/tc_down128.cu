// tcgen05 kernel for the WIDE 'down' block of the as-shipped 512-model: PreActFixupResBlock 'down'
// 64 -> 128 channels (vq_ae/layers/conv_block.py:196-216; the last level of the n_down = 4 pyramid,
// model.py:144-148).  tc_down.cu keeps all weights and all four parity planes of every operand in
// shared memory, which does not fit at this width (240 KB of weights alone); here
//
//   * the prologue scatters the 16 x 32 input pixels of a tile into the four parity planes of
//     A1 = f16(elu(x + b1a) + b1b) and As = f16(x + b1c)                  (2 x 66 KB, as in tc_down.cu)
//   * the planes are then processed ONE AT A TIME through a single 33 KB U buffer:
//         G1: D1 = A1[pl] . W1^T              E1: U = f16(elu(D1 + b2a) + b2b)
//         G2: D2 += U . W2[pl]^T              Gs: D3 += As[pl] . Ws[pl]^T
//     (the commit behind G1 of plane pl + 1 also covers G2 / Gs of plane pl, so when E1 may start
//     writing U again the previous plane's reads are complete)
//   * V = f16(elu(D2 + b3a) + b3b) reuses the U buffer;  G3: D3 += V . (scale W3)^T;  out = D3 + (b4 + b1d)
//   * the thirteen weight matrices of a tile (W1, W2[pl], Ws[pl] per plane, then W3; 288 KB) stream
//     from L2 through a two-slot ring of 32 KB bulk copies with full / empty mbarriers
//
// TMEM: D1 | D2 | D3, 128 columns each.  Warps 0-15: prologue / epilogues (4 lane quarters x 4 column
// groups of 32), warp 16: MMA issue, warp 17: weight producer.  Same arithmetic as tc_down.cu
// (fp16 operands, fp32 accumulation, fast ELU).
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace vqae {
namespace {

using namespace tc;

constexpr int DW_OH = 8, DW_OW = 16;           // output tile = one M = 128 MMA tile
constexpr int DW_CI = 64, DW_CO = 128;
constexpr int DW_PLANE = 129;                  // pixel pitch between parity planes (odd)
constexpr int DW_PIX = 4 * DW_PLANE + 1;       // 517
constexpr uint32_t DW_LBO = DW_PIX * 16;       // A1 / As: [k-chunk][517 px][16 B]
constexpr uint32_t DW_ULBO = DW_PLANE * 16;    // U / V:   [k-chunk][129 px][16 B]
constexpr int DW_KCI = DW_CI / 8, DW_KCO = DW_CO / 8;
constexpr int DW_NW = 16, DW_WORKERS = DW_NW * 32, DW_THREADS = DW_WORKERS + 64;
constexpr int DW_NC = 32, DW_UCH = DW_NC / 8;
constexpr uint32_t DW_WI_LBO = DW_CO * 16, DW_WI_BYTES = DW_KCI * DW_WI_LBO;   // [128 x 64]: 16 KB
constexpr uint32_t DW_WO_LBO = DW_CO * 16, DW_WO_BYTES = DW_KCO * DW_WO_LBO;   // [128 x 128]: 32 KB
constexpr int DW_RING = 2, DW_ITEMS = 13;      // ring slots; weight matrices per tile
constexpr uint32_t DW_OFF_A1 = 0;
constexpr uint32_t DW_OFF_AS = DW_OFF_A1 + DW_KCI * DW_LBO;
constexpr uint32_t DW_OFF_U = DW_OFF_AS + DW_KCI * DW_LBO;
constexpr uint32_t DW_OFF_RING = DW_OFF_U + DW_KCO * DW_ULBO;
constexpr uint32_t DW_OFF_BAR = DW_OFF_RING + DW_RING * DW_WO_BYTES;
constexpr uint32_t DW_SMEM = DW_OFF_BAR + 64;
static_assert(DW_SMEM <= 227 * 1024, "shared memory");

// CTA barrier of the worker warps and the MMA warp (17 warps); the weight producer runs free of it --
// it blocks on ring slots that are released by MMAs of later phases
__device__ __forceinline__ void dw_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(DW_WORKERS + 32) : "memory");
}

struct Down128Args {
    const float* x;               // NHWC fp32 [B,H,W,64]
    float* out;                   // NHWC fp32 [B,H/2,W/2,128]
    const uint8_t* w;             // VQAE_PACK_DOWN_F16, c_in = 64: [W1 | W2 x4 | scale*W3 | Ws x4]
    int n_tiles, H, W, tiles_x, tiles_per_img;
    float b1a, b1b, b2a, b2b, b3a, b3b, b1c, bsum;
};

// item i of a tile's weight sequence -> byte offset in the packed weights and byte count
__device__ __forceinline__ void dw_item(int i, uint32_t& off, uint32_t& bytes) {
    constexpr uint32_t O_W2 = DW_WI_BYTES, O_W3 = O_W2 + 4 * DW_WO_BYTES, O_WS = O_W3 + DW_WO_BYTES;
    if (i == 12) { off = O_W3; bytes = DW_WO_BYTES; return; }
    const int pl = i / 3, k = i - 3 * pl;
    if (k == 0) { off = 0; bytes = DW_WI_BYTES; }
    else if (k == 1) { off = O_W2 + pl * DW_WO_BYTES; bytes = DW_WO_BYTES; }
    else { off = O_WS + pl * DW_WI_BYTES; bytes = DW_WI_BYTES; }
}

__global__ void __launch_bounds__(DW_THREADS, 1)
down128_tc_kernel(Down128Args a) {
    constexpr int MMA_WARP = DW_NW, PROD_WARP = DW_NW + 1;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_mma = sbase + DW_OFF_BAR;
    const uint32_t bar_full = bar_mma + 8, bar_empty = bar_full + 8 * DW_RING;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + DW_OFF_BAR + 8 + 16 * DW_RING);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t leader = lane == 0;
    const int my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_items = my_tiles * DW_ITEMS;

    if (tid == 0) {
        mbar_init(bar_mma, 1);
        for (int s = 0; s < DW_RING; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        fence_mbar_init();
    }
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), 512);
    for (int i = tid; i < (int)(DW_OFF_RING / 16); i += DW_THREADS)
        *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t idesc = make_idesc_bf16(128, DW_CO);
    const uint32_t tD1 = tmem_base, tD2 = tmem_base + DW_CO, tD3 = tmem_base + 2 * DW_CO;

    const uint64_t dA1 = make_desc(sbase + DW_OFF_A1, DW_LBO, 128);
    const uint64_t dAS = make_desc(sbase + DW_OFF_AS, DW_LBO, 128);
    const uint64_t dU = make_desc(sbase + DW_OFF_U, DW_ULBO, 128);
    const uint64_t dR = make_desc(sbase + DW_OFF_RING, DW_WO_LBO, 128);      // W_LBO is the same for both shapes

    // ---- weight producer: one lane streams the item sequence of all of this CTA's tiles ----
    if (warp == PROD_WARP) {
        if (lane == 0) {
            for (int n = 0; n < total_items; ++n) {
                const int slot = n % DW_RING;
                if (n >= DW_RING) mbar_wait(bar_empty + 8 * slot, ((n / DW_RING) - 1) & 1);
                uint32_t off, bytes;
                dw_item(n % DW_ITEMS, off, bytes);
                mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
                bulk_g2s(sbase + DW_OFF_RING + slot * DW_WO_BYTES, a.w + off, bytes, bar_full + 8 * slot);
            }
        }
        __syncwarp();
    }
    int items_used = 0;                       // consumer state (MMA warp)
    // wait for the next weight matrix; returns its descriptor
    auto next_w = [&]() -> uint64_t {
        const int slot = items_used % DW_RING;
        mbar_wait(bar_full + 8 * slot, (items_used / DW_RING) & 1);
        tc_fence_after_sync();
        return dR + (uint64_t)((slot * DW_WO_BYTES) >> 4);
    };
    auto release_w = [&]() {                  // slot is free once the MMAs issued so far have completed
        umma_commit(bar_empty + 8 * (items_used % DW_RING), leader);
        ++items_used;
    };

    const int q4 = warp & 3, grp = (warp >> 2) & 3;
    const int row_in_tile = q4 * 32 + lane;
    const uint32_t t_lane = (uint32_t)(q4 * 32) << 16;
    const uint32_t t_col = grp * DW_NC;
    const int kc0 = grp * DW_UCH;
    constexpr int F4 = DW_NC / 4, SROW = DW_NC + 4;
    float* stage = reinterpret_cast<float*>(smem + DW_OFF_A1) + (warp < DW_NW ? warp : 0) * 32 * SROW;
    uint32_t mma_phase = 0;
    const int Ho = a.H / 2, Wo = a.W / 2;

    for (int tile = blockIdx.x; warp != PROD_WARP && tile < a.n_tiles; tile += gridDim.x) {
        const int img = tile / a.tiles_per_img;
        const int trem = tile - img * a.tiles_per_img;
        const int r0 = (trem / a.tiles_x) * DW_OH, c0 = (trem % a.tiles_x) * DW_OW;
        const float* ximg = a.x + (size_t)img * a.H * a.W * DW_CI;
        float* oimg = a.out + (size_t)img * Ho * Wo * DW_CO;

        // ---- P: A1 / As of the 16 x 32 input pixels, scattered into the four parity planes ----
        if (warp < DW_NW) {
            constexpr int ITEMS = 4 * 128 * DW_KCI;
            constexpr int PB = 4;
            for (int base = tid; base < ITEMS; base += DW_WORKERS * PB) {
                float vv[PB][8];
                int dst[PB];
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const int id = base + u * DW_WORKERS;
                    const int ip = id / DW_KCI, kc = id - ip * DW_KCI;
                    const int iy = ip >> 5, ix = ip & 31;
                    StreamIO<float>::load8(ximg + ((size_t)(2 * r0 + iy) * a.W + 2 * c0 + ix) * DW_CI + kc * 8, vv[u]);
                    const int m = ((iy & 1) * 2 + (ix & 1)) * DW_PLANE + (iy >> 1) * DW_OW + (ix >> 1);
                    dst[u] = kc * (int)DW_LBO + m * 16;
                }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    *reinterpret_cast<uint4*>(smem + DW_OFF_A1 + dst[u]) = act_pack8(vv[u], a.b1a, a.b1b);
                    *reinterpret_cast<uint4*>(smem + DW_OFF_AS + dst[u]) = add_pack8(vv[u], a.b1c);
                }
            }
        }
        fence_proxy_async_smem();
        dw_sync();

        // G1 of plane 0
        if (warp == MMA_WARP) {
            tc_fence_after_sync();
            const uint64_t dW = next_w();
#pragma unroll
            for (int ks = 0; ks < DW_CI / 16; ++ks)
                umma_bf16(tD1, dA1 + (uint64_t)((ks * 2 * DW_LBO) >> 4),
                          dW + (uint64_t)((ks * 2 * DW_WI_LBO) >> 4), idesc, ks > 0, leader);
            release_w();
            umma_commit(bar_mma, leader);
            __syncwarp();
        }
#pragma unroll 1
        for (int pl = 0; pl < 4; ++pl) {
            // ---- E1: U = f16(elu(D1 + b2a) + b2b)  (D1 complete => the previous plane's G2 / Gs too)
            if (warp < DW_NW) {
                mbar_wait(bar_mma, mma_phase);
                tc_fence_after_sync();
                float v[DW_NC];
                tmem_ld<DW_NC>(tD1 + t_lane + t_col, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < DW_UCH; ++j)
                    *reinterpret_cast<uint4*>(smem + DW_OFF_U + (kc0 + j) * DW_ULBO + row_in_tile * 16) =
                        act_pack8(v + 8 * j, a.b2a, a.b2b);
                tc_fence_before_sync();
            }
            mma_phase ^= 1;
            fence_proxy_async_smem();
            dw_sync();
            // ---- G2 += U . W2[pl]^T;  Gs: D3 += As[pl] . Ws[pl]^T;  G1 of the next plane ----
            if (warp == MMA_WARP) {
                tc_fence_after_sync();
                {
                    const uint64_t dW = next_w();
#pragma unroll
                    for (int ks = 0; ks < DW_CO / 16; ++ks)
                        umma_bf16(tD2, dU + (uint64_t)((ks * 2 * DW_ULBO) >> 4),
                                  dW + (uint64_t)((ks * 2 * DW_WO_LBO) >> 4), idesc, (pl | ks) > 0, leader);
                    release_w();
                }
                {
                    const uint64_t dW = next_w();
#pragma unroll
                    for (int ks = 0; ks < DW_CI / 16; ++ks)
                        umma_bf16(tD3, dAS + (uint64_t)((pl * DW_PLANE * 16 + ks * 2 * DW_LBO) >> 4),
                                  dW + (uint64_t)((ks * 2 * DW_WI_LBO) >> 4), idesc, (pl | ks) > 0, leader);
                    release_w();
                }
                if (pl < 3) {
                    const uint64_t dW = next_w();
#pragma unroll
                    for (int ks = 0; ks < DW_CI / 16; ++ks)
                        umma_bf16(tD1, dA1 + (uint64_t)(((pl + 1) * DW_PLANE * 16 + ks * 2 * DW_LBO) >> 4),
                                  dW + (uint64_t)((ks * 2 * DW_WI_LBO) >> 4), idesc, ks > 0, leader);
                    release_w();
                }
                umma_commit(bar_mma, leader);
                __syncwarp();
            }
        }
        // ---- E2: V = f16(elu(D2 + b3a) + b3b) over U ----
        if (warp < DW_NW) {
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            float v[DW_NC];
            tmem_ld<DW_NC>(tD2 + t_lane + t_col, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < DW_UCH; ++j)
                *reinterpret_cast<uint4*>(smem + DW_OFF_U + (kc0 + j) * DW_ULBO + row_in_tile * 16) =
                    act_pack8(v + 8 * j, a.b3a, a.b3b);
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        fence_proxy_async_smem();
        dw_sync();
        // ---- G3: D3 += V . (scale W3)^T ----
        if (warp == MMA_WARP) {
            tc_fence_after_sync();
            const uint64_t dW = next_w();
#pragma unroll
            for (int ks = 0; ks < DW_CO / 16; ++ks)
                umma_bf16(tD3, dU + (uint64_t)((ks * 2 * DW_ULBO) >> 4),
                          dW + (uint64_t)((ks * 2 * DW_WO_LBO) >> 4), idesc, 1u, leader);
            release_w();
            umma_commit(bar_mma, leader);
            __syncwarp();
        }
        // ---- E3: out = D3 + (b4 + b1d), transposed through shared memory (the A1 region is dead) ----
        if (warp < DW_NW) {
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            float v[DW_NC];
            tmem_ld<DW_NC>(tD3 + t_lane + t_col, v);
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
                const int p = q4 * 32 + rr;
                const size_t off = ((size_t)(r0 + (p >> 4)) * Wo + c0 + (p & 15)) * DW_CO + kc0 * 8 + c4 * 4;
                float4 d = *reinterpret_cast<const float4*>(stage + rr * SROW + 4 * c4);
                d.x += a.bsum; d.y += a.bsum; d.z += a.bsum; d.w += a.bsum;
                *reinterpret_cast<float4*>(oimg + off) = d;
            }
            __syncwarp();
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        dw_sync();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool down128_tc_supported(int H, int W, int CI) {
    return CI == DW_CI && H >= 2 * DW_OH && W >= 2 * DW_OW && H % (2 * DW_OH) == 0 && W % (2 * DW_OW) == 0;
}

int down128_tc(const float* x, float* out, const void* w_packed, const float* scalars8, int64_t B,
               int H, int W, int CI, int sm_count, cudaStream_t stream) {
    if (!x || !out || !w_packed || !scalars8 || B <= 0) return VQAE_ERR_BAD_ARG;
    if (!down128_tc_supported(H, W, CI)) return VQAE_ERR_UNSUPPORTED;
    Down128Args a;
    a.x = x; a.out = out; a.w = reinterpret_cast<const uint8_t*>(w_packed);
    a.H = H; a.W = W;
    a.tiles_x = (W / 2) / DW_OW; a.tiles_per_img = ((H / 2) / DW_OH) * a.tiles_x;
    const int64_t nt = B * a.tiles_per_img;
    if (nt > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    a.n_tiles = (int)nt;
    a.b1a = scalars8[0]; a.b1b = scalars8[1]; a.b2a = scalars8[2]; a.b2b = scalars8[3];
    a.b3a = scalars8[4]; a.b3b = scalars8[5]; a.b1c = scalars8[6]; a.bsum = scalars8[7];
    auto kern = down128_tc_kernel;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SMEM));
        attr_set.cur() = true;
    }
    const int grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;
    kern<<<grid, DW_THREADS, DW_SMEM, stream>>>(a);
    return check_launch();
}

}  // namespace vqae
