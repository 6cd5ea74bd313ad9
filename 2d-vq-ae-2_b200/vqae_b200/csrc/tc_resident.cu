// Image-resident tcgen05 kernel for a RUN of consecutive PreActFixupResBlocks in mode 'same' whose
// image a thread-block cluster can own: the 50-block trunks (vq_ae/model.py:150-153,240-263) with the
// adjacent post / pre layers of the pyramids (conv_block.py:18-91) at C = 64 or 128, 32 x 32, and the
// five-block runs at C = 32, 64 x 64; block arithmetic conv_block.py:196-216.  (Written up for the
// C = 64 shape; RsCfg below lists how the others map onto it.)
//
// The fp32 residual stream never leaves the SM: a cluster of four CTAs owns one image per *slot*,
// CTA r holding rows 8r .. 8r+7 (256 pixels = two 128-lane M-tiles) of the residual in TENSOR MEMORY
// (2 x 64 fp32 columns).  Per block and slot:
//     P   workers: tcgen05.ld residual -> A1 = bf16(elu(x + b1a) + b1b) -> shared operand buffer
//     G1  tcgen05.mma  D  = A1 . W1^T                      (2 M-tiles x 4 k-steps)
//     E1  workers: U = bf16(elu(D + b2a) + b2b) -> operand buffer, wrap-around columns duplicated;
//         a pusher warp then copies the first / last row into the neighbour CTAs' halo rows through
//         distributed shared memory (st.async, bytes counted on the receiver's mbarrier)
//     G2  nine taps accumulate into D; a tap is a constant start-address shift of the A descriptor;
//         the three dy = 0 taps need no halo row and go first
//     E2  workers: V = bf16(elu(D + b3a) + b3b) -> operand buffer
//     G3  tcgen05.mma  R += V . (scale W3)^T               accumulates straight into the residual
// so the residual add costs nothing and there is no global-memory traffic between the first load and
// the last store of an image (bias4 is carried as a running scalar and folded into the next b1a).
// Pixels of a tile are stored COLUMN-major with a 10-row pitch (8 rows + 2 halo rows): eight
// consecutive rows of one column are one 8-row core-matrix group, 16 columns at a constant stride
// (SBO = 160 B) are one M-tile, so an 8 x 32 tile is exactly two M-tiles -- no padded pixels are
// multiplied -- and tap (dy, dx) is the shift (dx * 10 + dy) * 16 B.
// Each CTA runs TWO slots (two different images, 2 x (128 residual + 128 accumulator) = 512 TMEM
// columns) half a block out of phase: while the nine taps of one slot occupy the tensor pipe, the 16
// worker warps run E2 -> P -> E1 of the other slot, half strip by half strip.  Every (slot, half strip)
// has its OWN MMA issue warp: one issuing thread only gets a 128 x 64 x 16 MMA every ~80 cycles out of
// the tensor pipe, several streams together reach 48 (profiles/mma_bench_issuers.py).  Each slot streams
// its weight matrices through its own ring of bulk copies, fed by its own producer warp.
// There is ONE operand buffer per slot (A1, U and V overwrite each other in place); a neighbour pushes
// the halo rows of block i+1 only after this CTA has signalled that the taps of block i have read the
// old rows (free barriers), which leaves 2 x 64 KB of shared memory for the weight rings.
// Warp roles: 0-15 workers | MMA issue warps | weight producers | halo pusher.  Every wait carries a
// watchdog (__trap after 2^26 polls): a protocol error faults instead of hanging the device.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace vqae {
namespace {

using namespace tc;

constexpr int RS_TH = 8;
constexpr int RS_NW = 16;                               // worker warps
constexpr int RS_SCAL_BLOCKS = 128;                     // runs up to this length keep their scalars in shared memory
constexpr int RS_NPUSH = 4;                             // halo pusher warps (a st.async holds its warp for
                                                        // about one DSMEM round trip: one warp alone needed
                                                        // ~1.7 k cycles for the 17 stores of a push)
constexpr int RS_PR = RS_TH + 2;                        // rows per stored column (with halo rows)
constexpr uint32_t RS_SBO = RS_PR * 16;                 // 8-row group stride = one column

// Shapes (C channels, W x W pixels; a cluster of W / 8 CTAs owns an image, one CTA a strip of 8 rows =
// W / 16 M-tiles, handled as two HALVES of MPH M-tiles each):
//   C = 64,  W = 32  (256-model trunks)          2 M-tiles, 2 slots, cluster 4
//   C = 128, W = 32  (512-model trunks)          2 M-tiles, 1 slot (256 + 256 TMEM columns), cluster 4
//   C = 32,  W = 64  (the five-block 'same' runs of the 256-model pyramids)
//                                                4 M-tiles, 2 slots, cluster 8
//   (C = 16, W = 128 instantiates and is correct with clusters of 16, but at five blocks per run the
//    image load / store and the few resident clusters leave it no faster than the tile kernel -- 1.03 vs
//    1.00 ms at batch 256 -- so it is not dispatched)
// In every case a slot's residual and accumulators take MT * C = 128 (256) TMEM columns each, an operand
// buffer is (W + 2) * 10 pixels * C * 2 B ~ 42 (85) KB, and a worker unit (one half, one phase) is
// 128 * MPH pixels x C channels = 8 192 (16 384) activations for the 16 worker warps.
template <int C, int W>
struct RsCfg {
    static_assert((C == 64 && W == 32) || (C == 128 && W == 32) || (C == 32 && W == 64) ||
                  (C == 16 && W == 128), "resident trunk kernel: unsupported shape");
    static constexpr int CL = W / RS_TH;                // cluster size (square images)
    static constexpr int MT = W / 16, MPH = MT / 2;     // M-tiles per strip / per half
    static constexpr int NSLOT = 2 * MT * C <= 256 ? 2 : 1;
    static constexpr int ISSUERS = 2 * NSLOT;           // one MMA issue warp per (slot, half)
    // Weights.  C <= 64: ONE buffer of 11 matrices per CTA, entry q = matrix q of the current block, used
    // by both slots (slot 1 half a block later) and refilled with the next block's matrix once all four
    // issue warps have released it: ring position and descriptor offset of every MMA are compile-time
    // constants (no R2UR traffic in front of the MMAs) and each block's weights are fetched once.
    // C = 128 (32 KB matrices): a 4-deep ring in issue order with a run-time position.
    static constexpr bool SHARED_W = C <= 64;
    static constexpr int RING = SHARED_W ? 11 : 4;
    static constexpr int NRING = SHARED_W ? 1 : NSLOT;  // rings (and producer warps) per CTA
    // + one weight producer warp per ring + the halo pusher warps
    static constexpr int THREADS = (RS_NW + ISSUERS + NRING + RS_NPUSH) * 32;
    static constexpr int NPIX = (W + 2) * RS_PR;        // stored pixels per operand buffer
    static constexpr uint32_t LBO = NPIX * 16;          // k-chunk (8 channels) stride
    static constexpr uint32_t BUF = (C / 8) * LBO;      // operand buffer of one slot
    static constexpr uint32_t WMAT = C * C * 2;
    static constexpr uint32_t WLBO = C * 16;
    static constexpr uint32_t HALO_BYTES = (W + 2) * C * 2;      // one halo row incl. wrap columns
    static constexpr uint32_t OFF_W = NSLOT * BUF;
    static constexpr uint32_t OFF_BAR = OFF_W + NRING * RING * WMAT;
    static constexpr uint32_t OFF_SCAL = OFF_BAR + 512;      // per-block scalars of up to RS_SCAL_BLOCKS blocks
    static constexpr uint32_t SMEM = OFF_SCAL + RS_SCAL_BLOCKS * 32;
    static constexpr int CPT = C * MPH / 4;             // TMEM columns (channels) per worker thread
    static_assert(CPT == 16 || CPT == 32, "16 or 32 channels per worker thread");
    static_assert(SMEM <= 232448, "shared memory budget");
};

struct ResidentArgs {
    const void* x;                // NHWC fp32 or fp16 [B,H,W,C]
    void* out;                    // NHWC, same element type (may alias x)
    int io_half;                  // 1: x / out are fp16 (the residual on chip is fp32 either way)
    const __nv_bfloat16* w;       // [n_blocks][11]: W1 | W2 tap 0..8 | scale*W3, each [k-chunk][n][8]
    const float* scal;            // [n_blocks][8] = b1a b1b b2a b2b b3a b3b b4 scale   (device)
    int n_blocks, n_img;
    long long* prof;              // optional [RS_PROF_HR][32] clock64 stamps of CTA 0 (profiling aid)
};
constexpr int RS_PROF_HR0 = 20, RS_PROF_HR = 8;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;"
                 ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cta_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done;
    uint32_t spins = 0;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 26)) __trap();
    } while (!done);
}
// 16 bytes into a peer CTA's shared memory through the async proxy (the proxy tcgen05.mma reads
// operands with), completing 16 transaction bytes on the peer's mbarrier: no fences on either side
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint4 v, uint32_t cluster_bar) {
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::
            "r"(cluster_addr),
        "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_bar)
        : "memory");
}
// wait with a watchdog: a protocol error traps instead of hanging the device
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity) {
    uint32_t done;
    uint32_t spins = 0;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 26)) __trap();
    } while (!done);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
        : "memory");
}
// 32 lanes x 8 columns (asynchronous until tcgen05.wait::ld)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                   "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Issue order of the nine taps: the three dy = 0 taps need no halo row and go first, so the pushes of
// the neighbours (DSMEM moves ~20 B/clk) overlap them; then dy = -1 (row from above), dy = +1 (below).
__host__ __device__ constexpr int tap_of(int i) { return i < 3 ? i + 3 : (i < 6 ? i - 3 : i); }

// The workers' static schedule.  Half-round hr (from -1): while slot a = hr & 1 runs the nine taps of
// its step ja = hr >> 1, the workers serve the other slot b: E2 of step jprev, then P and E1 of step
// jprev + 1 (slot 1 lags slot 0 by half a block).
struct HalfRound {
    int a, b, ja, jprev;
    bool g2, g3, g1;
};
__device__ __forceinline__ HalfRound half_round(int hr, int T0, int T1) {
    HalfRound h;
    h.a = hr & 1;
    h.b = h.a ^ 1;
    h.ja = hr >> 1;
    h.jprev = h.a ? h.ja : h.ja - 1;
    const int Ta = h.a ? T1 : T0, Tb = h.b ? T1 : T0;
    h.g2 = h.ja >= 0 && h.ja < Ta;
    h.g3 = h.jprev >= 0 && h.jprev < Tb;
    h.g1 = h.jprev + 1 < Tb;
    return h;
}

template <int C, int W>
__global__ void __cluster_dims__(RsCfg<C, W>::CL, 1, 1) __launch_bounds__(RsCfg<C, W>::THREADS, 1)
trunk_resident_tc_kernel(ResidentArgs a) {
    using Cfg = RsCfg<C, W>;
    constexpr int RS_C = C, RS_W = W, RS_H = W, RS_CL = Cfg::CL, MPH = Cfg::MPH;
    constexpr int RS_RING = Cfg::RING, RS_ISSUERS = Cfg::ISSUERS, NSLOT = Cfg::NSLOT, NRING = Cfg::NRING;
    constexpr bool SHARED_W = Cfg::SHARED_W;
    constexpr int CPT = Cfg::CPT, HC = CPT / 2;          // columns per thread / per half load
    constexpr uint32_t RS_LBO = Cfg::LBO, RS_BUF = Cfg::BUF, RS_WMAT = Cfg::WMAT, RS_WLBO = Cfg::WLBO;
    constexpr uint32_t RS_HALO_BYTES = Cfg::HALO_BYTES, RS_OFF_W = Cfg::OFF_W, RS_OFF_BAR = Cfg::OFF_BAR;
    // TMEM per slot: [R half 0 | R half 1 | D half 0 | D half 1], a half = MPH M-tiles of C columns
    constexpr uint32_t HALF_COLS = MPH * C, ACC_COL = 2 * HALF_COLS, SLOT_COLS = 2 * ACC_COL;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sW = sbase + RS_OFF_W;
    const uint32_t bar0 = sbase + RS_OFF_BAR;
    // barrier map (8 B each):
    //   acc[slot][m]  (4)  MMA -> workers: G1 / G2 / G3 of M-tile m complete (tcgen05.commit)
    //   wrk[slot][m]  (4)  workers -> MMA: A1 / V of M-tile m written (16 warps)
    //   u[slot]       (2)  workers -> MMA: U of BOTH M-tiles written (the taps read across them)
    //   halo[slot][dir] (4) row data from a neighbour landed (dir 0 = from the CTA above)
    //   free[slot][dir] (4) the neighbour's taps have consumed my last push (dir 0 = the CTA above)
    //   pushed[slot]  (2)  the pusher warp has read rows 0 / 7 of U: E2 may overwrite them with V
    //   full[NRING][RING] | empty[NRING][RING]   weight ring(s)
    const uint32_t bar_acc = bar0, bar_wrk = bar0 + 32, bar_u = bar0 + 64, bar_halo = bar0 + 80;
    const uint32_t bar_free = bar0 + 112, bar_pushed = bar0 + 144;
    const uint32_t bar_full = bar0 + 160, bar_empty = bar_full + 8 * NRING * RS_RING;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + RS_OFF_BAR + 160 + 16 * NRING * RS_RING);
    static_assert(160 + 16 * NRING * RS_RING + 4 <= 512, "barrier region");

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t leader = lane == 0;
    const uint32_t rank = cluster_ctarank();
    // Persistent clusters: cluster k works through the image groups k, k + NCL, k + 2 NCL, ... (a group
    // = one image per slot); the step counters of a slot simply run on across its images, so the next
    // image's blocks follow the previous one's without re-launching the cluster (a re-launch costs
    // ~20 us: all CTAs of the cluster must be free at once, TMEM allocation, barrier set-up, the
    // half-block lead-in of slot 1 -- 40 % of the time of a five-block run).
    const int cl_id = blockIdx.x / RS_CL, ncl = gridDim.x / RS_CL;
    const int n = a.n_blocks;
    const int n_groups = (a.n_img + NSLOT - 1) / NSLOT;
    const int my_groups = cl_id < n_groups ? (n_groups - cl_id + ncl - 1) / ncl : 0;
    // images of slot s: NSLOT * (cl_id + t * ncl) + s for t < cnt_s (the very last group may lack slot 1)
    const int cnt0 = my_groups;
    const int cnt1 = NSLOT == 2 ? my_groups - ((my_groups > 0 && NSLOT * (cl_id + (my_groups - 1) * ncl) + 1 >= a.n_img) ? 1 : 0) : 0;
    const int T0 = n * cnt0, T1 = n * cnt1;
    auto image_of = [&](int slot, int step) { return NSLOT * (cl_id + (step / n) * ncl) + slot; };
    const int hr_last = 2 * (T0 > T1 ? T0 : T1);

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(bar_acc + 8 * i, 1);
            mbar_init(bar_wrk + 8 * i, RS_NW);
            mbar_init(bar_halo + 8 * i, 1);
            mbar_init(bar_free + 8 * i, 2);              // both issue warps of the neighbour's slot
        }
        mbar_init(bar_u, 2 * RS_NW);
        mbar_init(bar_u + 8, 2 * RS_NW);
        mbar_init(bar_pushed, RS_NPUSH);
        mbar_init(bar_pushed + 8, RS_NPUSH);
        for (int s = 0; s < NRING * RS_RING; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            // every issue warp that reads the entry: both of a slot, of all active slots if shared
            mbar_init(bar_empty + 8 * s, SHARED_W ? 2 * NSLOT : 2);
        }
        fence_mbar_init();
    }
    if (warp == RS_NW) tmem_alloc(smem_u32(tmem_slot), 512);
    // the Fixup scalars of every block: a global load in front of a tcgen05 fence costs the workers an
    // L2 round trip (~800 cycles) per half-round, a shared-memory read does not
    float* scal_s = reinterpret_cast<float*>(smem + Cfg::OFF_SCAL);
    for (int i = tid; i < 8 * (a.n_blocks < RS_SCAL_BLOCKS ? a.n_blocks : RS_SCAL_BLOCKS); i += Cfg::THREADS)
        scal_s[i] = __ldg(a.scal + i);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    cluster_sync_all();                                  // peers' barriers are initialised
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    constexpr uint32_t idesc = make_idesc_bf16(128, C);

    if (warp >= RS_NW + RS_ISSUERS + NRING) {
        // ---------------- halo pushers: RS_NPUSH warps, off the workers' critical path --------------
        // After E1 of (slot b, step j1) they copy stored rows 1 and 8 of the slot's U buffer (all W + 2
        // stored columns, i.e. including the wrap-around duplicates E1 wrote) into stored row 9 of the
        // CTA above and stored row 0 of the CTA below with st.async: the data travel in the async proxy
        // the MMA reads operands with and complete bytes on the receiver's halo barrier (dir 0 = from
        // above), so neither side needs a fence.  A st.async keeps its warp busy for about a DSMEM round
        // trip (~200 cycles), so the 2 x 272 16-byte pieces of a push are spread over several warps; the
        // neighbours' "old rows consumed" barriers are waited for BEFORE U is complete (they have long
        // fired by then, and a cluster-scope acquire wait costs 400-800 cycles).
        const int pw = warp - (RS_NW + RS_ISSUERS + NRING);
        const uint32_t up_rank = (rank + RS_CL - 1) % RS_CL, dn_rank = (rank + 1) % RS_CL;
        const uint32_t up_base = mapa_u32(sbase, up_rank), dn_base = mapa_u32(sbase, dn_rank);
        const uint32_t up_bar = mapa_u32(bar_halo + 8, up_rank);     // its "from below" barrier
        const uint32_t dn_bar = mapa_u32(bar_halo, dn_rank);         // its "from above" barrier
        constexpr int PIECES = (RS_W + 2) * (RS_C / 8);              // 16-byte pieces per row
        for (int hr = -1; hr <= hr_last; ++hr) {
            const HalfRound h = half_round(hr, T0, T1);
            if (!h.g1) continue;
            const int b = h.b, j1 = h.jprev + 1;
            const bool pf = a.prof && blockIdx.x == 0 && lane == 0 && pw == 0 && b == 0 &&
                            j1 >= RS_PROF_HR0 / 2 && j1 < (RS_PROF_HR0 + RS_PROF_HR) / 2;
            long long* pp = a.prof + (j1 - RS_PROF_HR0 / 2) * 64;
            if (j1 > 0) {                                // the neighbours' taps of the previous block
                mbar_wait_cluster(bar_free + 16 * b, (j1 - 1) & 1);          // have read the old rows
                mbar_wait_cluster(bar_free + 16 * b + 8, (j1 - 1) & 1);
            }
            if (pf) pp[19] = clock64();
            mbar_wait_wd(bar_u + 8 * b, j1 & 1);         // both halves of U written (E1)
            if (pf) pp[18] = clock64();
            const uint32_t buf = (uint32_t)b * RS_BUF;
            for (int p = pw * 32 + lane; p < PIECES; p += 32 * RS_NPUSH) {
                const int kc = p / (RS_W + 2), cs = p - kc * (RS_W + 2);     // k-chunk, stored column
                const uint32_t off = buf + kc * RS_LBO + (uint32_t)(cs * RS_PR) * 16;
                const uint4 top = *reinterpret_cast<const uint4*>(smem + off + 1 * 16);       // my row 0
                const uint4 bot = *reinterpret_cast<const uint4*>(smem + off + RS_TH * 16);   // my row 7
                st_async_v4(up_base + off + (RS_TH + 1) * 16, top, up_bar + 16 * b);
                st_async_v4(dn_base + off, bot, dn_bar + 16 * b);
            }
            __syncwarp();
            if (pf) pp[21] = clock64();
            if (lane == 0) mbar_arrive(bar_pushed + 8 * b);
        }
    } else if (warp >= RS_NW + RS_ISSUERS) {
        // ---------------- weight producer(s): one warp (one thread) per ring ------------------------
        // Ring order = the issue order of a step: W1 | taps (tap_of) | scale * W3.  A producer blocks on
        // its own ring only (mbarrier.try_wait suspends the warp in hardware); a polling loop over two
        // rings with __nanosleep delivered the last taps of every block ~1 k cycles late.
        const int sl = warp - (RS_NW + RS_ISSUERS);
        if (lane == 0) {
            const int tot = (SHARED_W ? (T0 > T1 ? T0 : T1) : (sl ? T1 : T0)) * 11;
            for (int c = 0; c < tot; ++c) {
                const int rs = c % RS_RING;
                if (c >= RS_RING) mbar_wait_wd(bar_empty + 8 * (sl * RS_RING + rs), ((c / RS_RING) - 1) & 1);
                const int blk = (c / 11) % n, q = c % 11;
                const int m = q == 0 ? 0 : (q == 10 ? 10 : 1 + tap_of(q - 1));
                const uint32_t bf = bar_full + 8 * (sl * RS_RING + rs);
                mbar_arrive_expect_tx(bf, RS_WMAT);
                bulk_g2s(sW + (sl * RS_RING + rs) * RS_WMAT,
                         reinterpret_cast<const uint8_t*>(a.w) + ((size_t)blk * 11 + m) * RS_WMAT, RS_WMAT, bf);
            }
        }
    } else if (warp >= RS_NW) {
        // ---------------- MMA issue warps: one per (slot, M-tile group) ---------------------------
        // A single issuing thread gets one 128 x 64 x 16 MMA per ~80 cycles out of the tensor pipe; two
        // or more streams reach 48 (the shared-memory operand rate), profiles/mma_bench_issuers.py.
        const int iw = warp - RS_NW;
        const int slot = iw >> 1, m = iw & 1;
        const int T = slot ? T1 : T0;
        const uint32_t R = slot * SLOT_COLS + m * HALF_COLS, D = R + ACC_COL;
        int wcnt = 0;
        uint32_t wrk_par = 0;
        const int ring = SHARED_W ? 0 : slot;
        const uint64_t dW = make_desc(sW + ring * RS_RING * RS_WMAT, RS_WLBO, 128);
        // A operand of this half's first M-tile: own pixels start at stored column 16 MPH m + 1, row 1;
        // the next M-tile is 16 columns further
        const uint32_t a_base = sbase + (uint32_t)slot * RS_BUF +
                                (uint32_t)(((16 * MPH * m + 1) * RS_PR + 1) * 16);
        constexpr uint32_t A_MT = 16 * RS_PR * 16;
        const uint32_t b_acc = bar_acc + 8 * iw, b_wrk = bar_wrk + 8 * iw, b_u = bar_u + 8 * slot;
        auto wait_wrk = [&]() {
            mbar_wait_wd(b_wrk, wrk_par);
            wrk_par ^= 1u;
            tc_fence_after_sync();
        };
        // one weight matrix against this M-tile: D(+)= A(shift) . W^T
        long long ring_wait = 0;
        // q: position of the matrix inside a step (0 = W1, 1 + i = i-th tap issued, 10 = W3).  With the
        // shared buffer q is also the ring position (a compile-time constant after unrolling) and the
        // phase is the step's parity; with per-slot rings the position runs on with wcnt.
        auto gemm = [&](int q, int j, int shift_px, uint32_t d_col, bool acc_first) {
            const int rs = SHARED_W ? q : wcnt % RS_RING;
            const uint32_t par = SHARED_W ? (uint32_t)(j & 1) : (uint32_t)((wcnt / RS_RING) & 1);
            const long long tw = a.prof ? clock64() : 0;
            mbar_wait_wd(bar_full + 8 * (ring * RS_RING + rs), par);
            if (a.prof) ring_wait += clock64() - tw;
            tc_fence_after_sync();
            const uint64_t dWm = dW + (uint64_t)((rs * RS_WMAT) >> 4);
#pragma unroll
            for (int mi = 0; mi < MPH; ++mi) {
                const uint64_t dA = make_desc(a_base + mi * A_MT + shift_px * 16, RS_LBO, RS_SBO);
#pragma unroll
                for (int ks = 0; ks < RS_C / 16; ++ks)
                    umma_bf16(tmem_base + d_col + mi * RS_C, dA + (uint64_t)((ks * 2 * RS_LBO) >> 4),
                              dWm + (uint64_t)((ks * 2 * RS_WLBO) >> 4), idesc,
                              (acc_first || ks > 0) ? 1u : 0u, leader);
            }
            umma_commit(bar_empty + 8 * (ring * RS_RING + rs), leader);
            ++wcnt;
        };
        const uint32_t hb = bar_halo + 16 * slot;
        // the neighbours' "my push has been consumed" barriers: I am the CTA below my upper neighbour
        // (its free[slot][1]) and the CTA above my lower neighbour (its free[slot][0])
        const uint32_t up_free = mapa_u32(bar_free + 16 * slot + 8, (rank + RS_CL - 1) % RS_CL);
        const uint32_t dn_free = mapa_u32(bar_free + 16 * slot, (rank + 1) % RS_CL);
        for (int j = 0; j < T; ++j) {
            const bool pf = a.prof && blockIdx.x == 0 && lane == 0 && iw == 0 && j >= RS_PROF_HR0 / 2 &&
                            j < (RS_PROF_HR0 + RS_PROF_HR) / 2;
            long long* pp = a.prof + (j - RS_PROF_HR0 / 2) * 64;
            // ---- G1: D = A1 . W1^T ----
            if (pf) pp[0] = clock64();
            wait_wrk();
            if (pf) pp[1] = clock64();
            gemm(0, j, 0, D, false);
            umma_commit(b_acc, leader);
            // ---- G2: nine taps; dy = 0 first, then the halo rows as they arrive ----
            const uint32_t hp = j & 1;
            if (m == 0 && leader) {                      // 34 pixels x 128 B from each neighbour
                mbar_arrive_expect_tx(hb, RS_HALO_BYTES);
                mbar_arrive_expect_tx(hb + 8, RS_HALO_BYTES);
            }
            if (pf) pp[2] = clock64();
            mbar_wait_wd(b_u, hp);                       // U of both M-tiles written (E1)
            tc_fence_after_sync();
            if (pf) pp[3] = clock64();
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                if (i == 3) mbar_wait_wd(hb, hp);        // halo row from the CTA above
                if (i == 6) mbar_wait_wd(hb + 8, hp);    // halo row from the CTA below
                const int t = tap_of(i);
                if (pf) pp[23 + i] = clock64();
                gemm(1 + i, j, (t % 3 - 1) * RS_PR + (t / 3 - 1), D, i > 0);
            }
            umma_commit(b_acc, leader);
            if (pf) pp[4] = clock64();
            // my taps have read the halo rows: hand them back to the neighbours for the next block
            // (phase 3j + 1 of acc[slot][m]: G1, G2, G3 commits per step; their free barriers count
            // both issue warps of the slot)
            mbar_wait_wd(b_acc, (3 * j + 1) & 1);
            if (pf) { pp[7] = clock64(); pp[20] = ring_wait; }
            if (leader && j + 1 < T) {
                mbar_arrive_remote(up_free);
                mbar_arrive_remote(dn_free);
            }
            __syncwarp();
            // ---- G3: R += V . (scale W3)^T ----
            wait_wrk();
            if (pf) pp[5] = clock64();
            gemm(10, j, 0, R, true);
            umma_commit(b_acc, leader);
            if (pf) pp[6] = clock64();
            __syncwarp();
        }
        // shared weight buffer: a slot with fewer steps than the other one (odd batch: the last image
        // group has no slot 1) keeps releasing the entries it no longer reads, so that the producer's
        // release count per entry stays 2 * NSLOT
        if constexpr (SHARED_W) {
            const int t_max = T0 > T1 ? T0 : T1;
            for (int j = T; j < t_max; ++j)
                for (int q = 0; q < RS_RING; ++q) {
                    mbar_wait_wd(bar_full + 8 * q, (uint32_t)(j & 1));
                    if (leader) mbar_arrive(bar_empty + 8 * q);
                    __syncwarp();
                }
        }
    } else {
        // ---------------- workers: per unit (slot, M-tile, phase) one pixel x 16 channels per thread --
        // All 16 warps work on ONE M-tile at a time, so the MMA of M-tile 0 runs under the workers'
        // pass over M-tile 1 and the G3 / G1 round trips disappear from the chain.
        // TMEM lane quarter; M-tile inside the half; CPT-column group
        const int q4 = warp & 3, mi = (warp >> 2) % MPH, cq = (warp >> 2) / MPH;
        const int col0 = 16 * mi + 4 * q4 + (lane >> 3), row = lane & 7;   // pixel inside half 0
        const uint32_t t_off = ((uint32_t)(q4 * 32) << 16) + mi * RS_C + cq * CPT;
        const uint32_t kc_off = (uint32_t)(cq * (CPT / 8)) * RS_LBO;
        constexpr uint32_t M_PIX = 16 * MPH * RS_PR * 16;    // byte offset of half 1's pixels
        const uint32_t pix_own = (uint32_t)((col0 + 1) * RS_PR + row + 1) * 16 + kc_off;
        // wrap-around duplicates (circular padding): column 0 (M-tile 0) -> stored column 33,
        // column 31 (M-tile 1) -> stored column 0
        const bool wrap0 = col0 == 0, wrap1 = col0 == 16 * MPH - 1;
        const uint32_t pix_wrap0 = (uint32_t)((RS_W + 1) * RS_PR + row + 1) * 16 + kc_off;
        const uint32_t pix_wrap1 = (uint32_t)(row + 1) * 16 + kc_off;
        const size_t g_pix = ((size_t)(rank * RS_TH + row) * RS_W + col0) * RS_C + cq * CPT;

        uint32_t acc_par = 0;                            // bit 2s+m: parity of acc[s][m] to wait for next
        float cum0 = 0.f, cum1 = 0.f;                    // running sum of bias4 per slot
        auto wait_acc = [&](int sm) {
            mbar_wait_wd(bar_acc + 8 * sm, (acc_par >> sm) & 1);
            acc_par ^= 1u << sm;
            tc_fence_after_sync();
        };
        // CPT columns in two loads: the second is in flight while the first half is activated
        // (TMEM reads run at 64 B/clk per SM sub-partition, comparable to the SFU time of a unit)
        auto load_lo = [&](uint32_t taddr, float* v) {
            if constexpr (HC == 8) tmem_ld8(taddr, v);
            else tmem_ld16(taddr, *reinterpret_cast<float(*)[16]>(v));
            tmem_ld_wait();
            if constexpr (HC == 8) tmem_ld8(taddr + HC, v + HC);
            else tmem_ld16(taddr + HC, *reinterpret_cast<float(*)[16]>(v + HC));
        };
        // store the bf16 chunks of the first / second half of a thread's channels
        auto store_half = [&](uint32_t dst, const float* v, int half, float pre, float post) {
#pragma unroll
            for (int k = 0; k < HC / 8; ++k)
                st_cta_v4(dst + (half * (HC / 8) + k) * RS_LBO, act_pack8(v + half * HC + 8 * k, pre, post));
        };
        auto signal = [&](uint32_t bar) {
            tc_fence_before_sync();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        };

        int bn0 = 0, bn1 = 0, bp0 = 0, bp1 = 0;        // per slot: block of the next step to prepare / of the previous one
        for (int hr = -1; hr <= hr_last; ++hr) {
            const HalfRound h = half_round(hr, T0, T1);
            const int b = h.b;
            const bool pf = a.prof && blockIdx.x == 0 && tid == 0 && hr >= RS_PROF_HR0 &&
                            hr < RS_PROF_HR0 + RS_PROF_HR;
            long long* pp = a.prof + (hr - RS_PROF_HR0) * 32;
            if (pf) pp[8] = clock64();
            const uint32_t buf = sbase + (uint32_t)b * RS_BUF;
            // the scalars of both blocks this half-round touches, fetched ahead of the first wait
            // block indices of the two steps this half-round touches, kept as running counters (a run-time
            // "% n" is a 30-instruction I2F / MUFU.RCP / F2I chain, twice, in front of the first wait of
            // every half-round)
            const int blk_n = b ? bn1 : bn0, blk_p = b ? bp1 : bp0;
            if (h.g1) {
                const int nx = blk_n + 1 == n ? 0 : blk_n + 1;
                if (b) { bp1 = blk_n; bn1 = nx; } else { bp0 = blk_n; bn0 = nx; }
            }
            const float* scal = n <= RS_SCAL_BLOCKS ? scal_s : a.scal;
            const float4 sp1 = *(reinterpret_cast<const float4*>(scal + blk_p * 8) + 1);
            const float4 sn0 = *reinterpret_cast<const float4*>(scal + blk_n * 8);
            if (h.g3) {
                // ---- E2: V over U (own pixels), M-tile by M-tile ----
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    // V overwrites U in place, and the taps of EACH half read one column of the other
                    // half (dx = +-1 across the seam): both halves' taps must be complete before the
                    // first store.  (Waiting per half let half 0's V land in column 16 MPH - 1 while
                    // half 1's last taps were still reading it -- seven wrong pixels in the seam
                    // column in ~1 % of launches once the tap phase got faster;
                    // profiles/determinism_diffpos.py.)
                    if (m == 0) {
                        wait_acc(2 * b);                 // nine taps of step jprev complete, half 0
                        wait_acc(2 * b + 1);             // ... and half 1
                    }
                    if (pf) pp[9 + m] = clock64();
                    float v[CPT];
                    load_lo(tmem_base + b * SLOT_COLS + ACC_COL + m * HALF_COLS + t_off, v);
                    // V goes over U in place: the pusher must have read rows 0 / 7 (long done)
                    if (m == 0) mbar_wait_wd(bar_pushed + 8 * b, h.jprev & 1);
                    const uint32_t dst = buf + pix_own + m * M_PIX;
                    store_half(dst, v, 0, sp1.x, sp1.y);
                    tmem_ld_wait();
                    store_half(dst, v, 1, sp1.x, sp1.y);
                    signal(bar_wrk + 8 * (2 * b + m));
                }
                if (pf) pp[11] = clock64();
            }
            // the residual holds x_{blk+1} - sum(bias4) once G3 has completed
            const float cum_prev = b ? cum1 : cum0;
            const float cum = h.g3 ? cum_prev + sp1.z : cum_prev;
            const bool last = h.g3 && blk_p == n - 1;
            if (h.g3 || h.g1) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const uint32_t Rm = tmem_base + b * SLOT_COLS + m * HALF_COLS + t_off;
                    float v[CPT];
                    if (h.g3) {
                        wait_acc(2 * b + m);             // G3 of step jprev complete
                        if (pf) pp[12 + m] = clock64();
                    }
                    if (last) {
                        const int img = image_of(b, h.jprev);
                        load_lo(Rm, v);
                        tmem_ld_wait();
                        const size_t eoff = (size_t)img * RS_H * RS_W * RS_C + g_pix + m * 16 * MPH * RS_C;
                        if (a.io_half) {
                            __half* o = reinterpret_cast<__half*>(a.out) + eoff;
#pragma unroll
                            for (int k = 0; k < CPT / 4; ++k)
                                StreamIO<__half>::store4(o + 4 * k, make_float4(v[4 * k] + cum, v[4 * k + 1] + cum,
                                                                                v[4 * k + 2] + cum, v[4 * k + 3] + cum));
                        } else {
                            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + eoff);
#pragma unroll
                            for (int k = 0; k < CPT / 4; ++k)
                                o[k] = make_float4(v[4 * k] + cum, v[4 * k + 1] + cum, v[4 * k + 2] + cum,
                                                   v[4 * k + 3] + cum);
                        }
                    }
                    if (h.g1) {
                        // ---- P: A1 from the residual (first block of an image: from global memory) ----
                        float pre = sn0.x;
                        if (blk_n == 0) {
                            const int img = image_of(b, h.jprev + 1);
                            const size_t eoff = (size_t)img * RS_H * RS_W * RS_C + g_pix + m * 16 * MPH * RS_C;
                            if (a.io_half) {
                                const __half* s2 = reinterpret_cast<const __half*>(a.x) + eoff;
#pragma unroll
                                for (int k = 0; k < CPT / 8; ++k) {
                                    float t8[8];
                                    StreamIO<__half>::load8(s2 + 8 * k, t8);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) v[8 * k + e] = t8[e];
                                }
                            } else {
                                const float4* s4 = reinterpret_cast<const float4*>(
                                    reinterpret_cast<const float*>(a.x) + eoff);
#pragma unroll
                                for (int k = 0; k < CPT / 4; ++k) {
                                    const float4 t = __ldg(s4 + k);
                                    v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
                                }
                            }
                            if constexpr (CPT == 16) tmem_st16(Rm, *reinterpret_cast<float(*)[16]>(v));
                            else tmem_st32(Rm, *reinterpret_cast<float(*)[32]>(v));
                            tmem_st_wait();
                        } else {
                            pre += cum;
                            load_lo(Rm, v);
                        }
                        const uint32_t dst = buf + pix_own + m * M_PIX;
                        store_half(dst, v, 0, pre, sn0.y);
                        tmem_ld_wait();
                        store_half(dst, v, 1, pre, sn0.y);
                        signal(bar_wrk + 8 * (2 * b + m));
                    }
                }
                if (pf) pp[14] = clock64();
            }
            if (b) cum1 = last ? 0.f : cum; else cum0 = last ? 0.f : cum;
            if (h.g1) {
                const int j1 = h.jprev + 1;
                // ---- E1: U over A1, wrap columns, halo rows into the neighbours ----
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    wait_acc(2 * b + m);                 // G1 of step j1 complete
                    if (pf) pp[15 + m] = clock64();
                    float v[CPT];
                    load_lo(tmem_base + b * SLOT_COLS + ACC_COL + m * HALF_COLS + t_off, v);
                    constexpr int NCH = CPT / 8;         // 16-byte chunks per thread
                    uint4 u[NCH];
                    const uint32_t dst = buf + pix_own + m * M_PIX;
                    const bool wrap = m ? wrap1 : wrap0;
                    const uint32_t dw = buf + (m ? pix_wrap1 : pix_wrap0);
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        if (k == NCH / 2) tmem_ld_wait();
                        u[k] = act_pack8(v + 8 * k, sn0.z, sn0.w);
                        st_cta_v4(dst + k * RS_LBO, u[k]);
                        if (wrap) st_cta_v4(dw + k * RS_LBO, u[k]);
                    }
                    signal(bar_u + 8 * b);
                }
                if (pf) pp[17] = clock64();
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                  // no CTA leaves while a peer may still push
    if (warp == RS_NW) tmem_dealloc(tmem_base, 512);
}

// clusters of this instantiation the device holds at once (0 if the query fails)
template <int C, int W>
int resident_clusters() {
    using Cfg = RsCfg<C, W>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(Cfg::CL * 64);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = Cfg::CL; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, trunk_resident_tc_kernel<C, W>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

template <int C, int W>
int launch_resident(const ResidentArgs& a, int64_t B, cudaStream_t stream) {
    using Cfg = RsCfg<C, W>;
    static PerDevice<int> max_clusters_dev{};       // 0 = not queried yet on this device
    int& max_clusters = max_clusters_dev.cur();
    if (max_clusters == 0) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(trunk_resident_tc_kernel<C, W>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        if (Cfg::CL > 8)                                 // clusters of 16 CTAs are an opt-in size
            VQAE_CUDA_TRY(cudaFuncSetAttribute(trunk_resident_tc_kernel<C, W>,
                                               cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        max_clusters = resident_clusters<C, W>();
    }
    // one image per slot and group; persistent: no more clusters than the device holds at once
    int64_t clusters = (B + Cfg::NSLOT - 1) / Cfg::NSLOT;
    if (max_clusters > 0 && clusters > max_clusters) clusters = max_clusters;
    if (clusters * Cfg::CL > 0x7fffffff) return VQAE_ERR_UNSUPPORTED;
    trunk_resident_tc_kernel<C, W><<<(unsigned)clusters * Cfg::CL, Cfg::THREADS, Cfg::SMEM, stream>>>(a);
    return check_launch();
}

}  // namespace

static long long* g_resident_prof = nullptr;
void trunk_resident_set_prof(long long* dev_ptr) { g_resident_prof = dev_ptr; }

bool trunk_resident_supported(int64_t B, int H, int W, int C) {
    if (B <= 0 || H != W) return false;
    return (W == 32 && (C == 64 || C == 128)) || (W == 64 && C == 32);
}

// resident clusters of the C = 64 kernel the device can hold at once (each runs one image pair at a time)
int trunk_resident_max_clusters(int* out) {
    VQAE_CUDA_TRY(cudaFuncSetAttribute(trunk_resident_tc_kernel<64, 32>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)RsCfg<64, 32>::SMEM));
    *out = resident_clusters<64, 32>();
    return *out > 0 ? VQAE_OK : VQAE_ERR_CUDA;
}

int trunk_resident_tc(const void* x, void* out, int io_dtype, const void* w_packed_all,
                      const float* scalars_dev, int n_blocks, int64_t B, int H, int W, int C,
                      cudaStream_t stream) {
    if (!x || !out || !w_packed_all || !scalars_dev || B <= 0 || n_blocks <= 0) return VQAE_ERR_BAD_ARG;
    if (io_dtype != VQAE_DT_F32 && io_dtype != VQAE_DT_F16) return VQAE_ERR_UNSUPPORTED;
    if (!trunk_resident_supported(B, H, W, C)) return VQAE_ERR_UNSUPPORTED;
    ResidentArgs a;
    a.x = x; a.out = out; a.io_half = io_dtype == VQAE_DT_F16;
    a.w = reinterpret_cast<const __nv_bfloat16*>(w_packed_all);
    a.scal = scalars_dev;
    a.n_blocks = n_blocks; a.n_img = (int)B;
    a.prof = g_resident_prof;
    switch (C) {
        case 64: return launch_resident<64, 32>(a, B, stream);
        case 128: return launch_resident<128, 32>(a, B, stream);
        case 32: return launch_resident<32, 64>(a, B, stream);
    }
    return VQAE_ERR_UNSUPPORTED;
}

}  // namespace vqae
