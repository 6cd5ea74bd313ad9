// tcgen05 (5th-gen tensor core) kernels: bf16 operands, fp32 accumulation in TMEM.
//
// same_block_tc_kernel: one whole PreActFixupResBlock in mode 'same' at C = 64, 32x32
// (reference: vq_ae/layers/conv_block.py:196-216; the trunk pre_enc_layers / post_enc_layers,
// vq_ae/model.py:150-153,240-263, and the post-'down' blocks of the last pyramid level) as three
// chained implicit GEMMs per half-image tile, all intermediates on chip:
//
//   A1 = bf16(elu(x + b1a) + b1b)          18 rows x 32 px (16 rows + circular y-halo), from HBM
//   D1 = A1 . W1^T                          tcgen05.mma  M=128 N=64 K=64           (1x1)
//   U  = bf16(elu(D1 + b2a) + b2b)          written "padded-linear": 18 rows x 34 px, the two
//                                           extra columns hold the circular x-halo
//   D2 = sum_taps U[q + dy*34 + dx] . W2^T  9 taps x K=64; a tap is a constant 16-byte-granular
//                                           shift of the operand start address           (3x3)
//   V  = bf16(elu(D2 + b3a) + b3b)
//   D3 = V . W3^T                                                                        (1x1)
//   out = x + scale * D3 + b4               fp32, to HBM
//
// Operands use the un-swizzled K-major canonical layout [k-chunk][pixel][8 x bf16], so any
// pixel shift keeps the descriptor regular (tc_common.cuh).  W1/W3 stay resident in shared
// memory, the nine 8 KB W2 taps stream through a 4-slot ring filled by bulk async copies
// (UBLKCP) from L2.  Warps 0-7: prologue/epilogue math; warp 8: MMA issue; warp 9: W2 producer.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace vqae {
namespace {

using namespace tc;

// -----------------------------------------------------------------------------------------------
// fused 'same' block on tcgen05, any H x W that tiles into 16 x 32 pixel tiles.
//   CP: channel count seen by the MMAs (16, 32, 64);  CR: real channels (8 runs zero-padded as 16)
// Tile = 16 rows x 32 cols of one image plus a 1-pixel circular halo ring: 18 x 34 = 612 pixels in
// "padded-linear" order q = lr * 34 + lc.  GEMM1 (1x1) is pointwise, so it is simply evaluated on
// all 612 (5 M-tiles of 128); a 3x3 tap is then the constant operand shift dy * 34 + dx.
// -----------------------------------------------------------------------------------------------
constexpr int SB_TH = 16, SB_TW = 32, SB_PW = SB_TW + 2;
constexpr int SB_NPAD = (SB_TH + 2) * SB_PW;      // 612 padded pixels
constexpr int SB_XPIX = 641;                       // operand region pitch: 5 x 128 px, odd
constexpr int SB_UPIX = 711;                       // 35 + 5*128 + 35 px, odd
constexpr uint32_t SB_XLBO = SB_XPIX * 16;
constexpr uint32_t SB_ULBO = SB_UPIX * 16;
constexpr int SB_RING = 4;

template <int CP, int CR>
struct SameCfg {
    static constexpr int KCH = CP / 8;             // 16-byte k-chunks per pixel seen by the MMA
    static constexpr int KCR = CR / 8;             // ... that hold real channels
    static constexpr bool RING = (CP == 64);       // W2 does not fit next to the operands: stream it
    static constexpr int NW = (CP == 64) ? 16 : (CP == 32 ? 8 : 4);   // worker warps
    static constexpr int WORKERS = NW * 32;
    static constexpr int THREADS = WORKERS + 64;   // + MMA warp + weight-producer warp
    static constexpr int NG = NW / 4;              // column groups (4 warps cover the 128 TMEM lanes)
    static constexpr int NC = CP / NG;             // TMEM columns per epilogue unit (16)
    static constexpr int UCH = (CR < CP) ? KCR : NC / 8;            // real k-chunks per unit
    static constexpr uint32_t WLBO = CP * 16;
    static constexpr uint32_t WMAT = KCH * WLBO;                    // one CP x CP bf16 matrix
    static constexpr uint32_t OFF_X = 0;
    static constexpr uint32_t OFF_U = OFF_X + KCH * SB_XLBO;
    static constexpr uint32_t OFF_W = OFF_U + KCH * SB_ULBO;        // RING: W1, W3, ring[4]; else all 11
    static constexpr uint32_t W_BYTES = (RING ? (2 + SB_RING) : 11) * WMAT;
    static constexpr uint32_t OFF_BAR = OFF_W + W_BYTES;
    static constexpr uint32_t SMEM = OFF_BAR + 128;
    static constexpr int TMEM_COLS = CP == 64 ? 512 : (CP == 32 ? 256 : 128);
    static constexpr int MIN_CTAS = CP == 64 ? 1 : (CP == 32 ? 2 : 4);
};

struct SameBlockArgs {
    const void* x;                // NHWC fp32 or fp16 [B,H,W,CR]
    void* out;                    // NHWC, same element type [B,H,W,CR]
    const __nv_bfloat16* w;       // 11 matrices [W1 | W2 tap 0..8 | W3], each [k-chunk][n][8]
    int n_tiles, H, W, tiles_x, tiles_per_img;
    float b1a, b1b, b2a, b2b, b3a, b3b, b4, scale;
    unsigned stagger_ns;          // start delay of CTAs with slack (0 = off)
    long long* prof;              // optional [gridDim.x][8] phase timestamps of each CTA's 1st tile
};


template <int CP, int CR, typename TIO>
__global__ void __launch_bounds__(SameCfg<CP, CR>::THREADS, SameCfg<CP, CR>::MIN_CTAS)
same_block_tc_kernel(SameBlockArgs a) {
    using IO = StreamIO<TIO>;
    const TIO* const xg = reinterpret_cast<const TIO*>(a.x);
    TIO* const og = reinterpret_cast<TIO*>(a.out);
    using Cfg = SameCfg<CP, CR>;
    constexpr int KCH = Cfg::KCH, KCR = Cfg::KCR, NW = Cfg::NW, NC = Cfg::NC, UCH = Cfg::UCH;
    constexpr uint32_t WLBO = Cfg::WLBO, WMAT = Cfg::WMAT;
    constexpr int MMA_WARP = NW, PROD_WARP = NW + 1;

    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sX = sbase + Cfg::OFF_X, sU = sbase + Cfg::OFF_U, sW = sbase + Cfg::OFF_W;
    const uint32_t sW1 = sW;
    const uint32_t sW3 = Cfg::RING ? sW + WMAT : sW + 10 * WMAT;
    const uint32_t sW2 = Cfg::RING ? sW + 2 * WMAT : sW + WMAT;      // ring base / tap 0
    const uint32_t bar_mma = sbase + Cfg::OFF_BAR;                   // MMA phase complete
    const uint32_t bar_full = bar_mma + 8;                           // [SB_RING] tap landed
    const uint32_t bar_empty = bar_full + 8 * SB_RING;               // [SB_RING] tap consumed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_BAR + 8 + 16 * SB_RING);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // provably warp-uniform
    const uint32_t leader = lane == 0;
    const int my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_taps = my_tiles * 9;

    // ---- one-time setup ----
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        for (int s = 0; s < SB_RING; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        fence_mbar_init();
    }
    if (warp == MMA_WARP) tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
    {
        const uint4* g = reinterpret_cast<const uint4*>(a.w);
        if (Cfg::RING) {   // resident W1 / W3 only
            for (int i = tid; i < (int)(WMAT / 16); i += Cfg::THREADS) {
                *reinterpret_cast<uint4*>(smem + Cfg::OFF_W + i * 16) = __ldg(g + i);
                *reinterpret_cast<uint4*>(smem + Cfg::OFF_W + WMAT + i * 16) =
                    __ldg(g + 10 * (WMAT / 16) + i);
            }
        } else {
            for (int i = tid; i < (int)(11 * WMAT / 16); i += Cfg::THREADS)
                *reinterpret_cast<uint4*>(smem + Cfg::OFF_W + i * 16) = __ldg(g + i);
        }
        // zero the operand regions once: slack rows and zero-padded channels stay finite / zero
        for (int i = tid; i < (int)(Cfg::OFF_W / 16); i += Cfg::THREADS)
            *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t idesc = make_idesc_bf16(128, CP);
    const uint8_t* w2g = reinterpret_cast<const uint8_t*>(a.w) + WMAT;

    int taps_issued = 0;                  // producer state (PROD_WARP lane 0, RING only)
    if (Cfg::RING && warp == PROD_WARP && lane == 0) {
        for (; taps_issued < SB_RING && taps_issued < total_taps; ++taps_issued) {
            const uint32_t fb = bar_full + 8 * (taps_issued % SB_RING);
            mbar_arrive_expect_tx(fb, WMAT);
            bulk_g2s(sW2 + (taps_issued % SB_RING) * WMAT, w2g + (taps_issued % 9) * WMAT, WMAT, fb);
        }
    }
    int taps_used = 0;                    // consumer state (MMA_WARP lane 0, RING only)
    uint32_t mma_phase = 0;

    // epilogue geometry of a worker thread
    const int q4 = warp & 3;              // TMEM lane quarter this warp may access
    const int grp = (warp >> 2) % Cfg::NG;        // column group handled by this warp
    const int row_in_tile = q4 * 32 + lane;
    const uint32_t t_lane = (uint32_t)(q4 * 32) << 16;
    const uint32_t t_col = grp * NC;
    const int kc0 = grp * (NC / 8);               // first k-chunk of this thread's unit
    // E3 staging (warp-private, in the U region once G3 has consumed V): [32 rows][NCR + 4] fp32
    constexpr int NCR = UCH * 8;                  // real channels per epilogue unit
    constexpr int F4 = NCR / 4;                   // float4 per staged row
    constexpr int SROW = NCR + 4;
    float* stage = reinterpret_cast<float*>(smem + Cfg::OFF_U) + (warp < NW ? warp : 0) * 32 * SROW;

    // descriptor bases; per-MMA offsets are compile-time constants added to the low word
    const uint64_t dX = make_desc(sX, SB_XLBO, 128);
    const uint64_t dU = make_desc(sU, SB_ULBO, 128);
    const uint64_t dW1 = make_desc(sW1, WLBO, 128);
    const uint64_t dW3 = make_desc(sW3, WLBO, 128);
    const uint64_t dW2 = make_desc(sW2, WLBO, 128);

    // P: A1 = bf16(elu(x + b1a) + b1b) on the 18 x 34 halo'd tile (circular wrap) -> X region.
    // Loads are issued in batches of PB items per thread before any use (memory-level parallelism).
    auto prologue = [&](int tile) {
        const int img = tile / a.tiles_per_img;
        const int trem = tile - img * a.tiles_per_img;
        const int r0 = (trem / a.tiles_x) * SB_TH, c0 = (trem % a.tiles_x) * SB_TW;
        const TIO* ximg = xg + (size_t)img * a.H * a.W * CR;
        constexpr int ITEMS = SB_NPAD * KCR;
        constexpr int PB = (NW >= 16) ? 3 : 6;
        for (int base = tid; base < ITEMS; base += Cfg::WORKERS * PB) {
            float vv[PB][8];
            int dst[PB];
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                const int id = base + u * Cfg::WORKERS;
                dst[u] = -1;
                if (id < ITEMS) {
                    const int q = id / KCR, kc = id - q * KCR;
                    const int lr = q / SB_PW, lc = q - lr * SB_PW;
                    int row = r0 - 1 + lr, col = c0 - 1 + lc;
                    row = row < 0 ? row + a.H : (row >= a.H ? row - a.H : row);
                    col = col < 0 ? col + a.W : (col >= a.W ? col - a.W : col);
                    IO::load8(ximg + ((size_t)row * a.W + col) * CR + kc * 8, vv[u]);
                    dst[u] = kc * (int)SB_XLBO + q * 16;
                }
            }
#pragma unroll
            for (int u = 0; u < PB; ++u) {
                if (dst[u] >= 0)
                    *reinterpret_cast<uint4*>(smem + Cfg::OFF_X + dst[u]) = act_pack8(vv[u], a.b1a, a.b1b);
            }
        }
    };

    int tile = blockIdx.x;
    // De-phase the CTAs: the tiles of all SMs otherwise march in lockstep and their memory phases
    // (operand gather, residual epilogue) hit L2/HBM in the same burst.  CTAs that own one tile
    // fewer than the busiest ones have a whole tile time of slack, so they start half a tile late.
    if (a.stagger_ns > 0 && my_tiles < (a.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x)
        __nanosleep(a.stagger_ns);
    const bool prof_cta = a.prof != nullptr && tid == 0;
    if (prof_cta) a.prof[(size_t)blockIdx.x * 8 + 0] = clock64();
    if (warp < NW && tile < a.n_tiles) prologue(tile);
    fence_proxy_async_smem();
    __syncthreads();
    bool first = true;

    for (; tile < a.n_tiles; tile += gridDim.x) {
        const int img = tile / a.tiles_per_img;
        const int trem = tile - img * a.tiles_per_img;
        const int r0 = (trem / a.tiles_x) * SB_TH, c0 = (trem % a.tiles_x) * SB_TW;
        const TIO* ximg = xg + (size_t)img * a.H * a.W * CR;
        TIO* oimg = og + (size_t)img * a.H * a.W * CR;
        const bool prof = prof_cta && first;
        if (prof) a.prof[(size_t)blockIdx.x * 8 + 1] = clock64();

        // ---- G1: D1 = A1 . W1^T on 5 M-tiles of padded-linear pixels ----
        if (warp == MMA_WARP) {          // whole warp, warp-uniform operands; lane 0 issues
            {
                tc_fence_after_sync();
#pragma unroll
                for (int t = 0; t < 5; ++t)
#pragma unroll
                    for (int ks = 0; ks < CP / 16; ++ks)
                        umma_bf16(tmem_base + t * CP,
                                  dX + (uint64_t)((t * 128 * 16 + ks * 2 * SB_XLBO) >> 4),
                                  dW1 + (uint64_t)((ks * 2 * WLBO) >> 4), idesc, ks > 0, leader);
                umma_commit(bar_mma, leader);
            }
            __syncwarp();
        }
        // ---- E1: U[q] = bf16(elu(D1[q] + b2a) + b2b), same padded-linear order ----
        if (warp < NW) {
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            if (prof) a.prof[(size_t)blockIdx.x * 8 + 2] = clock64();
            for (int t = 0; t < 5; ++t) {
                float v[NC];
                tmem_ld<NC>(tmem_base + t_lane + t * CP + t_col, v);
                tmem_ld_wait();
                const int q = t * 128 + row_in_tile;
#pragma unroll
                for (int j = 0; j < UCH; ++j)
                    *reinterpret_cast<uint4*>(smem + Cfg::OFF_U + (kc0 + j) * SB_ULBO + q * 16) =
                        act_pack8(v + 8 * j, a.b2a, a.b2b);
            }
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        fence_proxy_async_smem();
        __syncthreads();
        if (prof) a.prof[(size_t)blockIdx.x * 8 + 3] = clock64();

        // ---- G2: D2[q] = sum over 9 taps of U[q + dy*34 + dx] . W2[tap]^T, q from 35;
        //      meanwhile the workers build the NEXT tile's A1 (the X region is free after G1) ----
        if (Cfg::RING && warp == PROD_WARP && lane == 0) {
            for (int i = 0; i < 9 && taps_issued < total_taps; ++i, ++taps_issued) {
                const int slot = taps_issued % SB_RING;
                mbar_wait(bar_empty + 8 * slot, ((taps_issued / SB_RING) - 1) & 1);
                mbar_arrive_expect_tx(bar_full + 8 * slot, WMAT);
                bulk_g2s(sW2 + slot * WMAT, w2g + (taps_issued % 9) * WMAT, WMAT,
                         bar_full + 8 * slot);
            }
        }
        if (warp == MMA_WARP) {          // whole warp, warp-uniform operands; lane 0 issues
            {
                tc_fence_after_sync();
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    uint64_t dW;
                    int slot = 0;
                    if (Cfg::RING) {
                        slot = taps_used % SB_RING;
                        mbar_wait(bar_full + 8 * slot, (taps_used / SB_RING) & 1);
                        tc_fence_after_sync();
                        dW = dW2 + (uint64_t)((slot * WMAT) >> 4);
                        ++taps_used;
                    } else {
                        dW = dW2 + (uint64_t)((tap * WMAT) >> 4);
                    }
                    const int shift = (tap / 3 - 1) * SB_PW + (tap % 3 - 1);
#pragma unroll
                    for (int t = 0; t < 5; ++t)
#pragma unroll
                        for (int ks = 0; ks < CP / 16; ++ks)
                            umma_bf16(tmem_base + t * CP,
                                      dU + (uint64_t)(((SB_PW + 1 + t * 128 + shift) * 16 + ks * 2 * SB_ULBO) >> 4),
                                      dW + (uint64_t)((ks * 2 * WLBO) >> 4), idesc, (tap | ks) > 0, leader);
                    if (Cfg::RING) umma_commit(bar_empty + 8 * slot, leader);
                }
                umma_commit(bar_mma, leader);
            }
            __syncwarp();
        }
        if (warp < NW) {
            if (tile + (int)gridDim.x < a.n_tiles) prologue(tile + gridDim.x);
            if (prof) a.prof[(size_t)blockIdx.x * 8 + 4] = clock64();
            // ---- E2: V[p] = bf16(elu(D2 + b3a) + b3b), 16 x 32 interior, pixel-linear, written
            //      over the U region (dead once G2 has completed) ----
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            if (prof) a.prof[(size_t)blockIdx.x * 8 + 5] = clock64();
            for (int t = 0; t < 5; ++t) {
                float v[NC];
                tmem_ld<NC>(tmem_base + t_lane + t * CP + t_col, v);
                tmem_ld_wait();
                const int q = SB_PW + 1 + t * 128 + row_in_tile;
                const int lr = q / SB_PW, pc = q - lr * SB_PW;
                if (lr <= SB_TH && pc >= 1 && pc <= SB_TW) {
                    const int p = (lr - 1) * SB_TW + pc - 1;
#pragma unroll
                    for (int j = 0; j < UCH; ++j)
                        *reinterpret_cast<uint4*>(smem + Cfg::OFF_U + (kc0 + j) * SB_ULBO + p * 16) =
                            act_pack8(v + 8 * j, a.b3a, a.b3b);
                }
            }
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        fence_proxy_async_smem();
        __syncthreads();

        // ---- G3: D3 = V . W3^T, 4 M-tiles ----
        if (warp == MMA_WARP) {          // whole warp, warp-uniform operands; lane 0 issues
            {
                tc_fence_after_sync();
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int ks = 0; ks < CP / 16; ++ks)
                        umma_bf16(tmem_base + t * CP,
                                  dU + (uint64_t)((t * 128 * 16 + ks * 2 * SB_ULBO) >> 4),
                                  dW3 + (uint64_t)((ks * 2 * WLBO) >> 4), idesc, ks > 0, leader);
                umma_commit(bar_mma, leader);
            }
            __syncwarp();
        }
        // ---- E3: out = x + scale * D3 + b4 (fp32).  Each warp transposes its 32 x NCR block
        //      through shared memory so that global loads/stores are 128-bit and line-coalesced;
        //      the residual rows of M-tile t+1 are requested before M-tile t is processed ----
        if (warp < NW) {
            const int rsub = lane / F4, c4 = lane % F4;
            auto x_off = [&](int t, int k) {
                const int p = t * 128 + q4 * 32 + rsub + k * (32 / F4);
                return ((size_t)(r0 + (p >> 5)) * a.W + c0 + (p & 31)) * CR + kc0 * 8 + c4 * 4;
            };
            float4 xr[F4], xn[F4];
#pragma unroll
            for (int k = 0; k < F4; ++k)
                xr[k] = IO::load4(ximg + x_off(0, k));
            mbar_wait(bar_mma, mma_phase);
            tc_fence_after_sync();
            if (prof) a.prof[(size_t)blockIdx.x * 8 + 6] = clock64();
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float v[NC];
                tmem_ld<NC>(tmem_base + t_lane + t * CP + t_col, v);
                if (t + 1 < 4) {
#pragma unroll
                    for (int k = 0; k < F4; ++k)
                        xn[k] = IO::load4(ximg + x_off(t + 1, k));
                }
                tmem_ld_wait();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < F4; ++j)
                    *reinterpret_cast<float4*>(stage + lane * SROW + 4 * j) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < F4; ++k) {
                    const int rr = rsub + k * (32 / F4);           // row within this warp's 32
                    const float4 d = *reinterpret_cast<const float4*>(stage + rr * SROW + 4 * c4);
                    float4 o;
                    o.x = fmaf(d.x, a.scale, a.b4) + xr[k].x;
                    o.y = fmaf(d.y, a.scale, a.b4) + xr[k].y;
                    o.z = fmaf(d.z, a.scale, a.b4) + xr[k].z;
                    o.w = fmaf(d.w, a.scale, a.b4) + xr[k].w;
                    IO::store4(oimg + x_off(t, k), o);
                }
#pragma unroll
                for (int k = 0; k < F4; ++k) xr[k] = xn[k];
            }
            tc_fence_before_sync();
        }
        mma_phase ^= 1;
        fence_proxy_async_smem();     // next tile's A1 (written during G2) -> async proxy
        __syncthreads();
        if (prof) a.prof[(size_t)blockIdx.x * 8 + 7] = clock64();
        first = false;
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int CP, int CR, typename TIO>
int launch_same_block(const SameBlockArgs& a, int sm_count, cudaStream_t stream) {
    using Cfg = SameCfg<CP, CR>;
    auto kern = same_block_tc_kernel<CP, CR, TIO>;
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

int same_block_tc(const void* x, void* out, int io_dtype, const void* w_packed,
                  const float* scalars8, int64_t B, int H, int W, int C, int sm_count,
                  long long* prof, cudaStream_t stream) {
    if (io_dtype != VQAE_DT_F32 && io_dtype != VQAE_DT_F16) return VQAE_ERR_UNSUPPORTED;
    if (!x || !out || !w_packed || !scalars8 || B <= 0) return VQAE_ERR_BAD_ARG;
    if (x == out) return VQAE_ERR_BAD_ARG;
    if (H < SB_TH || W < SB_TW || H % SB_TH != 0 || W % SB_TW != 0) return VQAE_ERR_UNSUPPORTED;
    SameBlockArgs a;
    a.x = x; a.out = out; a.w = reinterpret_cast<const __nv_bfloat16*>(w_packed);
    a.H = H; a.W = W; a.tiles_x = W / SB_TW; a.tiles_per_img = (H / SB_TH) * a.tiles_x;
    const int64_t nt = B * a.tiles_per_img;
    if (nt > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    a.n_tiles = (int)nt;
    a.b1a = scalars8[0]; a.b1b = scalars8[1]; a.b2a = scalars8[2]; a.b2b = scalars8[3];
    a.b3a = scalars8[4]; a.b3b = scalars8[5]; a.b4 = scalars8[6]; a.scale = scalars8[7];
    a.prof = prof;
    a.stagger_ns = (C == 64) ? 9000u : 0u;
    const bool h = io_dtype == VQAE_DT_F16;
    switch (C) {
        case 64: return h ? launch_same_block<64, 64, __half>(a, sm_count, stream)
                          : launch_same_block<64, 64, float>(a, sm_count, stream);
        case 32: return h ? launch_same_block<32, 32, __half>(a, sm_count, stream)
                          : launch_same_block<32, 32, float>(a, sm_count, stream);
        case 16: return h ? launch_same_block<16, 16, __half>(a, sm_count, stream)
                          : launch_same_block<16, 16, float>(a, sm_count, stream);
        case 8: return h ? launch_same_block<16, 8, __half>(a, sm_count, stream)
                         : launch_same_block<16, 8, float>(a, sm_count, stream);
    }
    return VQAE_ERR_UNSUPPORTED;
}

}  // namespace vqae
