// Fused projected quantiser on tcgen05: HBM-bound form of quantize.cu for NHWC latents, C = 64 or 128
// channels, fp32 / bf16 / fp16 input and output (the distance arithmetic is fp32 in every case,
// like torch.cdist under autocast -- SURVEY.md section 0, finding 3).
//
// Reference: ProjectedEMAVectorQuantizer2d.forward / EMAVectorQuantizer.forward in eval mode
// (vq_ae/layers/vq.py:96-154,185-192).  Per tile of 128 latent vectors, all on chip:
//
//   x tile   <- two TMA tensor copies (128 rows x 32 channels each, SWIZZLE_128B so that the
//               row-per-thread 128-bit reads below are bank-conflict-free), 2-stage ring [copy warp]
//   z        =  proj_in(x), fp32 fma chain in channel order (packed FFMA2), weights broadcast
//               from shared memory                                                   [proj warps]
//   A        =  bf16 hi/lo split of (z, z^2, z^3) per dimension                       [proj warps]
//   D        =  A . B^T  on tcgen05 (M=128, N=256 codes, K=80), B = hi/lo split of
//               (-4 e^3, 6 e^2, -4 e, sum e^4): D_k = sum_d (z_d - e_kd)^4 - sum_d z_d^4
//               up to |err| <= c * T,  T = sum_d (|z_d| + max_k |e_kd|)^4              [MMA warp]
//   idx      =  exact fp32 argmin (the arithmetic of quantize.cu, lowest index wins) over the
//               candidates { k : D_k <= min D + margin(T) } read back from TMEM        [epi warps]
//   out      =  E'[idx] gathered from shared memory, loss partial, near-tie count      [epi warps]
//
// The tensor-core product is only a *filter*: every reported index, loss term and near-tie flag
// comes from the same fp32 evaluation as the CUDA-core kernel, so results are bit-identical to it
// provided the filter keeps the true best and every code within the near-tie gap of it -- which
// the margin 4*c*T + 2*gap*T guarantees (validated by
// tests/test_gpu_quantize_tc.py through the diagnostic output).  Rows whose candidate list is
// empty (NaN/Inf) or longer than 8 fall back to the full exact scan inside the kernel.
#include <cuda.h>      // CUtensorMap (types only; the encoder is fetched through the runtime)
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace vqae {
namespace {

using namespace tc;

constexpr int TQ_M = 128;              // vectors per tile
constexpr int TQ_K = 256;              // codes
constexpr int TQ_D = 8;                // distance-space dimension
constexpr int TQ_KK = 80;              // contraction length: 24 features x 3 split products + 3 + pad
constexpr int TQ_CH = TQ_KK / 8;       // 16-byte k-chunks
constexpr int TQ_XS = 2;               // x stages
// warps 0-3 proj, 4 MMA issue, 5 copy, 6-7 idle, 8-15 epilogue (two groups).  The epilogue is the
// longest stage; the hardware arbiter favours the highest warp ids, so it gets them.
constexpr int TQ_THREADS = 512;
constexpr int TQ_W_PROJ = 0, TQ_W_MMA = 4, TQ_W_COPY = 5, TQ_W_EPI = 8;
constexpr float TQ_ERR_C = 1.0f / 32768.0f;     // c = 2^-15: bound on |D_k - exact| / T

template <int C, typename XT>
struct TqCfg {
    static constexpr int XB = sizeof(XT);                     // bytes per x / out element
    static constexpr int CPS = 128 / XB;                      // channels per 128-byte slab
    static constexpr int XH = C / CPS;                        // 128-byte column slabs per row
    static constexpr uint32_t X_SLAB = TQ_M * 128;            // one TMA box: 128 rows x 128 bytes
    static constexpr uint32_t X_STAGE = XH * X_SLAB;
    static constexpr uint32_t OFF_X = 0;
    static constexpr uint32_t OFF_TAB = OFF_X + TQ_XS * X_STAGE;
    // E' table in the OUTPUT element type; it lives in shared memory when that fits next to the
    // x stages (<= 64 KB: every cell but C = 128 fp32, whose rows are gathered through L1/L2)
    static constexpr bool TAB_SMEM = TQ_K * C * XB <= 64 * 1024;
    static constexpr bool TAB_CONVERT = XB != 4;              // fp32 table -> XT while loading
    static constexpr uint32_t TAB_BYTES = TAB_SMEM ? TQ_K * C * XB : 0;
    static constexpr uint32_t OFF_B = OFF_TAB + TAB_BYTES;
    static constexpr uint32_t B_LBO = TQ_K * 16;
    static constexpr uint32_t OFF_A = OFF_B + TQ_CH * B_LBO;
    static constexpr uint32_t A_LBO = TQ_M * 16;
    static constexpr uint32_t OFF_Z = OFF_A + TQ_CH * A_LBO;
    static constexpr uint32_t ZPITCH = 48;                    // z[8], T, pad
    static constexpr uint32_t OFF_CB = OFF_Z + 2 * TQ_M * ZPITCH;
    static constexpr uint32_t CBP = 12;                        // codebook row pitch in floats (48 B:
                                                               // random-row 128-bit reads spread over all banks)
    static constexpr uint32_t OFF_EMAX = OFF_CB + TQ_K * CBP * 4;
    static constexpr uint32_t OFF_W = OFF_EMAX + 32;          // proj_in weight [c][8] + bias[8]
    static constexpr uint32_t OFF_BAR = OFF_W + (C + 1) * TQ_D * 4;
    static constexpr uint32_t SMEM = OFF_BAR + 160 + 1024;      // + slack to align the base to 1024 B
};

template <int C, typename XT>
struct TqArgs {
    const XT* x;             // [N][C]
    XT* out;                 // [N][C] or null
    int64_t* idx;            // [N]
    float* partial;          // [gridDim.x] per-CTA sums of |z - e|^2
    uint32_t* tie_partial;   // [gridDim.x] per-CTA near-tie counts
    unsigned int* done_counter;   // zero on entry, zero again on exit
    float* loss;             // 1
    double inv_count;        // 1 / (N * 8)
    float commitment_cost;
    uint32_t* near_ties;     // or null
    float* z_out;            // [N][8] or null
    float* diag;             // [N][4] = {min D, T, #candidates, slow-path flag} or null
    long long* prof;         // [4][16] clock64 phase stamps of CTA 0, tiles 10..13 (or null)
    const float* embed;      // [256][8]
    const float* table;      // [256][C]
    int64_t N;
    int num_tiles;
    float tie_rel_gap;
    float margin;            // 4 c + 2 gap
    const float* w_in;       // proj_in weight [d][c] (device)
    const float* b_in;       // [d]
};

__device__ __forceinline__ void tmem_ld_wait32(float (&v)[32]) {
    // the "+f" operands tie the loaded registers to the wait, so no use can be scheduled above it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]),
                   "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]),
                   "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]),
                   "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]),
                   "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]),
                   "+f"(v[30]), "+f"(v[31])
                 :
                 : "memory");
}

// TMA tensor copy global -> shared (2-D tile), completion counted on an mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tmap, int c0,
                                            int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

// bit j = (v[j] < thr), two instructions per column: FADD (the sign of v - thr) and a funnel shift
// that pushes the sign bit into one of four byte-wide chains (columns 8a+7 .. 8a are pushed in
// that order, so bit i of chain a is column 8a + i).  NaN columns give a canonical NaN (sign 0,
// no hit); thr = +Inf hits everything; thr = NaN hits nothing (the caller then scans every code).
__device__ __forceinline__ uint32_t hit_mask32(const float (&v)[32], float thr) {
    uint32_t m[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int t = 7; t >= 0; --t)
#pragma unroll
        for (int a4 = 0; a4 < 4; ++a4)
            m[a4] = __funnelshift_l(__float_as_uint(v[a4 * 8 + t] - thr), m[a4], 1);
    return __byte_perm(__byte_perm(m[0], m[1], 0x0040), __byte_perm(m[2], m[3], 0x0040), 0x5410);
}

__device__ __forceinline__ float bf16_hi_as_float(float v) {      // value of bf16_rn(v)
    return __bfloat162float(__float2bfloat16_rn(v));
}
// 8 fp32 -> hi = bf16(v), lo = bf16(v - hi), packed in k order
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    float h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        h[i] = bf16_hi_as_float(v[i]);
        l[i] = v[i] - h[i];
    }
    hi = make_uint4(pack_true_bf16(h[0], h[1]), pack_true_bf16(h[2], h[3]), pack_true_bf16(h[4], h[5]),
                    pack_true_bf16(h[6], h[7]));
    lo = make_uint4(pack_true_bf16(l[0], l[1]), pack_true_bf16(l[2], l[3]), pack_true_bf16(l[4], l[5]),
                    pack_true_bf16(l[6], l[7]));
}

#define TQ_PROF(slot)                                                                   \
    do {                                                                                \
        if (a.prof != nullptr && blockIdx.x == 0 && lane == 0 && it >= 10 && it < 14)   \
            a.prof[(it - 10) * 16 + (slot)] = clock64();                                \
    } while (0)

// 16 consecutive channels of one x row (128-byte-swizzled stage) as fp32
template <typename XT>
__device__ __forceinline__ void load_x16(const uint8_t* xr, uint32_t slab_bytes, int q, int sw,
                                         float (&v)[16]) {
    if constexpr (sizeof(XT) == 4) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = q * 4 + t;                          // 16-byte chunk of the row
            const float4 f = *reinterpret_cast<const float4*>(xr + (j >> 3) * slab_bytes +
                                                              (((j & 7) << 4) ^ sw));
            v[4 * t] = f.x; v[4 * t + 1] = f.y; v[4 * t + 2] = f.z; v[4 * t + 3] = f.w;
        }
    } else {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int j = q * 2 + t;
            const uint4 u = *reinterpret_cast<const uint4*>(xr + (j >> 3) * slab_bytes +
                                                            (((j & 7) << 4) ^ sw));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 f;
                if constexpr (std::is_same<XT, __half>::value)
                    f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                else
                    f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
                v[8 * t + 2 * i] = f.x; v[8 * t + 2 * i + 1] = f.y;
            }
        }
    }
}

template <typename XT>
__device__ __forceinline__ XT from_float(float v) {
    if constexpr (std::is_same<XT, __half>::value) return __float2half_rn(v);
    else if constexpr (std::is_same<XT, __nv_bfloat16>::value) return __float2bfloat16_rn(v);
    else return v;
}

template <int C, typename XT>
__global__ void __launch_bounds__(TQ_THREADS, 1)
quantize_tc_kernel(const __grid_constant__ TqArgs<C, XT> a, const __grid_constant__ CUtensorMap tmap) {
    using Cfg = TqCfg<C, XT>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B destinations must sit on the 1024-byte swizzle pattern
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + Cfg::OFF_BAR;
    const uint32_t bar_full_x = bar0;            // [2] x stage landed          (1 arrival + tx)
    const uint32_t bar_empty_x = bar0 + 16;      // [2] x stage read            (128 arrivals)
    const uint32_t bar_a_full = bar0 + 32;       //     A + z written           (128)
    const uint32_t bar_a_free = bar0 + 40;       //     MMA has read A          (commit)
    const uint32_t bar_acc_full = bar0 + 48;     // [2] accumulator ready       (commit)
    const uint32_t bar_acc_free = bar0 + 64;     // [2] accumulator + z read    (128)
    const uint32_t bar_table = bar0 + 88;        //     E' table landed         (1 arrival + tx)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_BAR + 80);
    float* red = reinterpret_cast<float*>(smem + Cfg::OFF_BAR + 96);      // [8]
    uint32_t* redt = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_BAR + 128);   // [8]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t leader = lane == 0;
    const int my_tiles = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    float* cb = reinterpret_cast<float*>(smem + Cfg::OFF_CB);
    float* emax = reinterpret_cast<float*>(smem + Cfg::OFF_EMAX);

    auto produce = [&](int it) {           // copy warp: one x tile into stage it & 1
        const int s = it & 1;
        const int64_t n0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TQ_M;
        if (lane == 0) {
            // rows past N are zero-filled by the TMA unit and still count towards the tx bytes
            mbar_arrive_expect_tx(bar_full_x + 8 * s, Cfg::X_STAGE);
#pragma unroll
            for (int h = 0; h < Cfg::XH; ++h)
                tma_load_2d(sbase + Cfg::OFF_X + s * Cfg::X_STAGE + h * Cfg::X_SLAB, &tmap,
                            h * Cfg::CPS, (int)n0, bar_full_x + 8 * s);
        }
        __syncwarp();
    };

    // ---- one-time setup ----
    if (tid == 0) {
        mbar_init(bar_full_x, 1);      mbar_init(bar_full_x + 8, 1);
        mbar_init(bar_empty_x, 64);    mbar_init(bar_empty_x + 8, 64);
        mbar_init(bar_a_full, 64);     mbar_init(bar_a_free, 1);
        mbar_init(bar_acc_full, 1);    mbar_init(bar_acc_full + 8, 1);
        mbar_init(bar_acc_free, 128);  mbar_init(bar_acc_free + 8, 128);
        mbar_init(bar_table, Cfg::TAB_CONVERT ? 64 : 1);
        fence_mbar_init();
    }
    if (warp == TQ_W_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
    __syncthreads();
    if (warp == TQ_W_COPY) {               // prime the HBM pipeline before anything else
        for (int it = 0; it < TQ_XS && it < my_tiles; ++it) produce(it);
    }
    for (int i = tid; i < TQ_K * TQ_D / 4; i += TQ_THREADS)
        *reinterpret_cast<float4*>(cb + (i >> 1) * Cfg::CBP + (i & 1) * 4) =
            __ldg(reinterpret_cast<const float4*>(a.embed) + i);
    if (Cfg::TAB_SMEM && !Cfg::TAB_CONVERT && a.out != nullptr && warp == TQ_W_COPY && lane == 0) {
        // E' table (64 KB) -> shared memory as 16 bulk async copies that complete on their own
        // barrier: the epilogue only needs it at its first output phase, so the load is off the
        // start-up path.  Every CTA reads the same bytes: each one starts at a different chunk,
        // otherwise all SMs sweep the same L2 lines at the same moment.
        constexpr uint32_t CH = TQ_K * C * 4 / 16;
        mbar_arrive_expect_tx(bar_table, TQ_K * C * 4);
        for (int i = 0; i < 16; ++i) {
            const uint32_t ch = (uint32_t)(i + blockIdx.x) & 15u;
            bulk_g2s(sbase + Cfg::OFF_TAB + ch * CH,
                     reinterpret_cast<const uint8_t*>(a.table) + ch * CH, CH, bar_table);
        }
    }
    {
        float* wsm = reinterpret_cast<float*>(smem + Cfg::OFF_W);       // [c][d], then bias
        for (int i = tid; i < C * TQ_D; i += TQ_THREADS)
            wsm[(i % C) * TQ_D + i / C] = __ldg(a.w_in + i);
        if (tid < TQ_D) wsm[C * TQ_D + tid] = __ldg(a.b_in + tid);
    }
    if (tid < TQ_M)                        // constant k-chunk of A: (1, 1, 1, 0, 0, 0, 0, 0)
        *reinterpret_cast<uint4*>(smem + Cfg::OFF_A + 9 * Cfg::A_LBO + tid * 16) =
            make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);
    __syncthreads();
    if (tid < TQ_K) {
        // B row of code k: features paired with (z, z^2, z^3): (-4 e^3, 6 e^2, -4 e), then sum e^4
        const float* e = cb + tid * Cfg::CBP;
        float f1[8], f2[8], f3[8], e4 = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const float e2 = e[d] * e[d];
            f1[d] = -4.f * (e2 * e[d]);
            f2[d] = 6.f * e2;
            f3[d] = -4.f * e[d];
            e4 = fmaf(e2, e2, e4);
        }
        uint4 h1, l1, h2, l2, h3, l3;
        split8(f1, h1, l1);
        split8(f2, h2, l2);
        split8(f3, h3, l3);
        uint8_t* brow = smem + Cfg::OFF_B + tid * 16;
        // A parts are [hi | hi | lo]; B parts [hi | lo | hi]: hi*hi + hi*lo + lo*hi
        *reinterpret_cast<uint4*>(brow + 0 * Cfg::B_LBO) = h1;
        *reinterpret_cast<uint4*>(brow + 1 * Cfg::B_LBO) = h2;
        *reinterpret_cast<uint4*>(brow + 2 * Cfg::B_LBO) = h3;
        *reinterpret_cast<uint4*>(brow + 3 * Cfg::B_LBO) = l1;
        *reinterpret_cast<uint4*>(brow + 4 * Cfg::B_LBO) = l2;
        *reinterpret_cast<uint4*>(brow + 5 * Cfg::B_LBO) = l3;
        *reinterpret_cast<uint4*>(brow + 6 * Cfg::B_LBO) = h1;
        *reinterpret_cast<uint4*>(brow + 7 * Cfg::B_LBO) = h2;
        *reinterpret_cast<uint4*>(brow + 8 * Cfg::B_LBO) = h3;
        const float c0 = bf16_hi_as_float(e4);
        const float c1 = bf16_hi_as_float(e4 - c0);
        const float c2 = (e4 - c0) - c1;
        *reinterpret_cast<uint4*>(brow + 9 * Cfg::B_LBO) =
            make_uint4(pack_true_bf16(c0, c1), pack_true_bf16(c2, 0.f), 0u, 0u);
    } else if (tid < TQ_K + TQ_D) {
        const int d = tid - TQ_K;
        float m = 0.f;
        for (int k = 0; k < TQ_K; ++k) m = fmaxf(m, fabsf(cb[k * Cfg::CBP + d]));
        emax[d] = m;
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    float sq_acc = 0.f;                    // epilogue threads: sum over own vectors of |z - e|^2

    if (Cfg::TAB_SMEM && Cfg::TAB_CONVERT && a.out != nullptr && (warp == 6 || warp == 7)) {
        // 16-bit outputs: the E' table is rounded to the output type once, on its way into shared
        // memory, by the two otherwise idle warps (off the start-up path of every other role)
        const int t64 = tid - 6 * 32;
        XT* tab = reinterpret_cast<XT*>(smem + Cfg::OFF_TAB);
        constexpr int NV = TQ_K * C / 4;               // float4 groups
        for (int i0 = t64; i0 < NV; i0 += 64 * 8) {
            float4 f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (i0 + u * 64 < NV) f[u] = __ldg(reinterpret_cast<const float4*>(a.table) + i0 + u * 64);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (i0 + u * 64 < NV) {
                    XT* d = tab + (size_t)(i0 + u * 64) * 4;
                    d[0] = from_float<XT>(f[u].x); d[1] = from_float<XT>(f[u].y);
                    d[2] = from_float<XT>(f[u].z); d[3] = from_float<XT>(f[u].w);
                }
        }
        mbar_arrive(bar_table);
    }

    if (warp == TQ_W_COPY) {
        // ================= copy warp =================
        for (int it = TQ_XS; it < my_tiles; ++it) {
            mbar_wait(bar_empty_x + 8 * (it & 1), ((it >> 1) - 1) & 1);
            produce(it);
        }
    } else if (warp == TQ_W_MMA) {
        // ================= MMA warp =================
        const uint32_t idesc = make_idesc_true_bf16(128, TQ_K);
        const uint64_t dA = make_desc(sbase + Cfg::OFF_A, Cfg::A_LBO, 128);
        const uint64_t dB = make_desc(sbase + Cfg::OFF_B, Cfg::B_LBO, 128);
        for (int it = 0; it < my_tiles; ++it) {
            const int b = it & 1;
            mbar_wait(bar_a_full, it & 1);
            if (it >= 2) mbar_wait(bar_acc_free + 8 * b, ((it >> 1) - 1) & 1);
            tc_fence_after_sync();
            TQ_PROF(0);
#pragma unroll
            for (int ks = 0; ks < TQ_KK / 16; ++ks)
                umma_bf16(tmem_base + b * TQ_K, dA + (uint64_t)((ks * 2 * Cfg::A_LBO) >> 4),
                          dB + (uint64_t)((ks * 2 * Cfg::B_LBO) >> 4), idesc, ks > 0, leader);
            umma_commit(bar_a_free, leader);
            umma_commit(bar_acc_full + 8 * b, leader);
            __syncwarp();
        }
    } else if (warp < TQ_W_PROJ + 4) {
        // ================= proj warps: z = proj_in(x), features, A operand =================
        // Two warp pairs take alternate tiles (pair p: x stage p, z/TMEM buffer p); every thread
        // owns two rows, so each broadcast weight load feeds 8 FFMA2 on 8 independent chains.
        const int pw = warp - TQ_W_PROJ, pair = pw >> 1;
        const int row0 = (pw & 1) * 64 + lane;                 // and row0 + 32
        float em[8], bi[8];
        const float4* wsm4 = reinterpret_cast<const float4*>(smem + Cfg::OFF_W);
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            em[d] = emax[d];
            bi[d] = reinterpret_cast<const float*>(smem + Cfg::OFF_W)[C * TQ_D + d];
        }
        const int s = pair, b = pair;
        // 128B swizzle: 16-byte chunk j of row r sits at chunk (j ^ (r & 7)) of its 128-byte line
        const uint8_t* xr = smem + Cfg::OFF_X + s * Cfg::X_STAGE + row0 * 128;
        const int sw = (row0 & 7) << 4;                        // same for row0 + 32
        for (int it = pair; it < my_tiles; it += 2) {
            const int64_t n0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TQ_M + row0;
            if ((pw & 1) == 0) TQ_PROF(1);
            mbar_wait(bar_full_x + 8 * s, (it >> 1) & 1);
            if ((pw & 1) == 0) TQ_PROF(2);
            // z[d] = sum_c x[c] * w[d][c] as packed fp32x2 FMAs (FFMA2): the same per-element fmaf
            // chain in channel order as quantize.cu, two output dims per instruction
            float2 zp[2][4];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int d = 0; d < 4; ++d) zp[r][d] = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < C / 16; ++q) {                 // 16 channels at a time
                float xsa[16], xsb[16];
                load_x16<XT>(xr, Cfg::X_SLAB, q, sw, xsa);
                load_x16<XT>(xr + 32 * 128, Cfg::X_SLAB, q, sw, xsb);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int c = q * 16 + u;
                    const float4 wa = wsm4[c * 2], wb = wsm4[c * 2 + 1];
                    const float2 w0 = make_float2(wa.x, wa.y), w1 = make_float2(wa.z, wa.w);
                    const float2 w2 = make_float2(wb.x, wb.y), w3 = make_float2(wb.z, wb.w);
                    const float2 xx = make_float2(xsa[u], xsa[u]);
                    const float2 yy = make_float2(xsb[u], xsb[u]);
                    zp[0][0] = __ffma2_rn(xx, w0, zp[0][0]);
                    zp[1][0] = __ffma2_rn(yy, w0, zp[1][0]);
                    zp[0][1] = __ffma2_rn(xx, w1, zp[0][1]);
                    zp[1][1] = __ffma2_rn(yy, w1, zp[1][1]);
                    zp[0][2] = __ffma2_rn(xx, w2, zp[0][2]);
                    zp[1][2] = __ffma2_rn(yy, w2, zp[1][2]);
                    zp[0][3] = __ffma2_rn(xx, w3, zp[0][3]);
                    zp[1][3] = __ffma2_rn(yy, w3, zp[1][3]);
                }
            }
            // hand the x stage back only after every value read from it has been consumed by the fmaf
            // chains above (the arrive used to sit right behind the last group's loads)
            fence_proxy_async_smem();
            mbar_arrive(bar_empty_x + 8 * s);
            if ((pw & 1) == 0) TQ_PROF(3);
            uint4 hh[2][3], ll[2][3];
            float zz[2][8], TT[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float* z = zz[r];
                z[0] = zp[r][0].x; z[1] = zp[r][0].y; z[2] = zp[r][1].x; z[3] = zp[r][1].y;
                z[4] = zp[r][2].x; z[5] = zp[r][2].y; z[6] = zp[r][3].x; z[7] = zp[r][3].y;
                const int64_t n = n0 + 32 * r;
                if (n >= a.N) {                            // tail rows: keep the MMA inputs finite
#pragma unroll
                    for (int d = 0; d < 8; ++d) z[d] = 0.f;
                }
#pragma unroll
                for (int d = 0; d < 8; ++d) z[d] += bi[d];
                float z2[8], z3[8], T = 0.f;
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    z2[d] = z[d] * z[d];
                    z3[d] = z2[d] * z[d];
                    float sd = fabsf(z[d]) + em[d];
                    sd *= sd;
                    T = fmaf(sd, sd, T);
                }
                TT[r] = T;
                split8(z, hh[r][0], ll[r][0]);
                split8(z2, hh[r][1], ll[r][1]);
                split8(z3, hh[r][2], ll[r][2]);
                if (a.z_out != nullptr && n < a.N) {
                    float4* zo = reinterpret_cast<float4*>(a.z_out + n * TQ_D);
                    zo[0] = make_float4(z[0], z[1], z[2], z[3]);
                    zo[1] = make_float4(z[4], z[5], z[6], z[7]);
                }
            }
            // z/T slot b: wait for the epilogue of tile it - 2; A is single-buffered: wait for the
            // previous tile's MMA.  ORDER MATTERS: this warp pair only looks at every other phase of
            // bar_a_free (the other pair takes the phases in between), and a parity wait cannot tell
            // phase it - 1 from phase it - 3.  If this pair ran ahead while the MMA of tile it - 2 was
            // still held up by a slow epilogue, the barrier would still be in phase it - 2, whose
            // parity differs from the one waited for, and the wait would fall through: A(it) could
            // then overwrite A(it - 1) before the MMA has read it.  The epilogue of tile it - 2 can
            // only have finished after MMA(it - 2), so after the first wait the barrier is in phase
            // it - 1 or later and the second wait means what it says.
            if (it >= 2) mbar_wait(bar_acc_free + 8 * b, ((it >> 1) - 1) & 1);
            if (it >= 1) mbar_wait(bar_a_free, (it - 1) & 1);
            if ((pw & 1) == 0) TQ_PROF(4);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = row0 + 32 * r;
                uint8_t* arow = smem + Cfg::OFF_A + row * 16;
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    *reinterpret_cast<uint4*>(arow + (0 + f) * Cfg::A_LBO) = hh[r][f];
                    *reinterpret_cast<uint4*>(arow + (3 + f) * Cfg::A_LBO) = hh[r][f];
                    *reinterpret_cast<uint4*>(arow + (6 + f) * Cfg::A_LBO) = ll[r][f];
                }
                float* zr =
                    reinterpret_cast<float*>(smem + Cfg::OFF_Z + (b * TQ_M + row) * Cfg::ZPITCH);
                const float* z = zz[r];
                *reinterpret_cast<float4*>(zr) = make_float4(z[0], z[1], z[2], z[3]);
                *reinterpret_cast<float4*>(zr + 4) = make_float4(z[4], z[5], z[6], z[7]);
                zr[8] = TT[r];
            }
            fence_proxy_async_smem();
            mbar_arrive(bar_a_full);
            if ((pw & 1) == 0) TQ_PROF(5);
        }
    } else if (warp >= TQ_W_EPI) {
        // ================= epilogue warps: candidate filter, exact argmin, outputs =============
        // two groups of four warps (TMEM lane quarter = warp & 3) take alternate tiles
        const int wq = warp & 3, b = (warp - TQ_W_EPI) >> 2;
        const int row = wq * 32 + lane;
        uint32_t ties = 0;
        bool table_ready = false;
        for (int it = b; it < my_tiles; it += 2) {
            const int64_t n0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * TQ_M;
            const int64_t n = n0 + row;
            const bool ok = n < a.N;
            if (wq == 0) TQ_PROF(6);
            mbar_wait(bar_acc_full + 8 * b, (it >> 1) & 1);
            tc_fence_after_sync();
            if (wq == 0) TQ_PROF(7);
            const float* zr =
                reinterpret_cast<const float*>(smem + Cfg::OFF_Z + (b * TQ_M + row) * Cfg::ZPITCH);
            const float4 za = *reinterpret_cast<const float4*>(zr);
            const float4 zb = *reinterpret_cast<const float4*>(zr + 4);
            const float T = zr[8];
            const float z[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + b * TQ_K;

            // pass 1: min over the 256 approximate distances
            float mn = INFINITY, mn1 = INFINITY, mn2 = INFINITY, mn3 = INFINITY;
            {
                float v0[32], v1[32];
                tmem_ld32(taddr, v0);
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    tmem_ld_wait32(v0);
                    tmem_ld32(taddr + (c + 1) * 32, v1);
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        mn = fminf(mn, fminf(v0[j], v0[j + 1]));
                        mn1 = fminf(mn1, fminf(v0[j + 2], v0[j + 3]));
                        mn2 = fminf(mn2, fminf(v0[j + 4], v0[j + 5]));
                        mn3 = fminf(mn3, fminf(v0[j + 6], v0[j + 7]));
                    }
                    tmem_ld_wait32(v1);
                    if (c + 2 < 8) tmem_ld32(taddr + (c + 2) * 32, v0);
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        mn = fminf(mn, fminf(v1[j], v1[j + 1]));
                        mn1 = fminf(mn1, fminf(v1[j + 2], v1[j + 3]));
                        mn2 = fminf(mn2, fminf(v1[j + 4], v1[j + 5]));
                        mn3 = fminf(mn3, fminf(v1[j + 6], v1[j + 7]));
                    }
                }
            }
            mn = fminf(fminf(mn, mn1), fminf(mn2, mn3));
            if (wq == 0) TQ_PROF(8);
            const float thr = fmaf(T, a.margin, mn);
            // pass 2: per-chunk hit masks of the columns within the margin; nothing else happens
            // while the TMEM buffer is held
            uint32_t hm[8];
            {
                float v0[32], v1[32];
                tmem_ld32(taddr, v0);
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    tmem_ld_wait32(v0);
                    tmem_ld32(taddr + (c + 1) * 32, v1);
                    hm[c] = hit_mask32(v0, thr);
                    tmem_ld_wait32(v1);
                    if (c + 2 < 8) tmem_ld32(taddr + (c + 2) * 32, v0);
                    hm[c + 1] = hit_mask32(v1, thr);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(bar_acc_free + 8 * b);         // TMEM buffer b and z slot b are free
            if (wq == 0) TQ_PROF(9);

            // exact fp32 evaluation of the hits in ascending k, identical arithmetic to
            // quantize_kernel (quantize.cu): d-ordered fmaf chain, strict '<' (lowest index wins).
            // The loop is warp-uniform and branch-free inside (trip count = most hits of any lane,
            // ~3-4); the masks stay in registers: `curm` holds the unvisited hits of chunk `cc`,
            // `nz` the chunks not yet opened.
            int cnt = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) cnt += __popc(hm[c]);
            const bool slow = cnt == 0;                // NaN/Inf row: scan every code
            uint32_t nz = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (slow) hm[c] = 0xffffffffu;
                nz |= (hm[c] != 0u) ? (1u << c) : 0u;
            }
            uint32_t curm = 0;
            int cc = 0;
            float best = INFINITY, second = INFINITY;
            int bidx = 0;
            auto pop = [&](bool& active) -> int {          // next unvisited hit of this lane
                if (curm == 0u && nz != 0u) {              // open the next non-empty chunk
                    cc = __ffs(nz) - 1;
                    nz &= nz - 1u;
                    const uint32_t s01 = (cc & 1) ? hm[1] : hm[0], s23 = (cc & 1) ? hm[3] : hm[2];
                    const uint32_t s45 = (cc & 1) ? hm[5] : hm[4], s67 = (cc & 1) ? hm[7] : hm[6];
                    const uint32_t s03 = (cc & 2) ? s23 : s01, s47 = (cc & 2) ? s67 : s45;
                    curm = (cc & 4) ? s47 : s03;
                }
                active = curm != 0u;
                const int k = active ? cc * 32 + __ffs(curm) - 1 : 0;
                curm &= curm - 1u;
                return k;
            };
            auto l4 = [&](int k, bool active) -> float {
                const float4 e0 = *reinterpret_cast<const float4*>(cb + k * Cfg::CBP);
                const float4 e1 = *reinterpret_cast<const float4*>(cb + k * Cfg::CBP + 4);
                float d, q, acc;
                d = z[0] - e0.x; q = d * d; acc = q * q;
                d = z[1] - e0.y; q = d * d; acc = fmaf(q, q, acc);
                d = z[2] - e0.z; q = d * d; acc = fmaf(q, q, acc);
                d = z[3] - e0.w; q = d * d; acc = fmaf(q, q, acc);
                d = z[4] - e1.x; q = d * d; acc = fmaf(q, q, acc);
                d = z[5] - e1.y; q = d * d; acc = fmaf(q, q, acc);
                d = z[6] - e1.z; q = d * d; acc = fmaf(q, q, acc);
                d = z[7] - e1.w; q = d * d; acc = fmaf(q, q, acc);
                return active ? acc : INFINITY;
            };
            auto take = [&](int k, float acc) {
                const bool lt = acc < best;                // false for NaN, like the scalar scan
                second = lt ? best : fminf(second, acc);
                best = lt ? acc : best;
                bidx = lt ? k : bidx;
            };
            while (__any_sync(0xffffffffu, (nz | curm) != 0u)) {
                // two hits per trip: two independent load + fmaf chains in flight
                bool act_a, act_b;
                const int ka = pop(act_a);
                const int kb = pop(act_b);
                const float da = l4(ka, act_a);
                const float db = l4(kb, act_b);
                take(ka, da);
                take(kb, db);
            }

            if (wq == 0) TQ_PROF(10);
            if (ok) {
                const float* e = cb + bidx * Cfg::CBP;
                float sq = 0.f;
#pragma unroll
                for (int d = 0; d < TQ_D; ++d) {
                    const float df = z[d] - e[d];
                    sq = fmaf(df, df, sq);
                }
                sq_acc += sq;
                a.idx[n] = bidx;
                if ((second - best) < a.tie_rel_gap * second) ++ties;
                if (a.diag != nullptr)
                    *reinterpret_cast<float4*>(a.diag + n * 4) =
                        make_float4(mn, T, (float)(slow ? TQ_K : cnt), slow ? 1.f : 0.f);
            }
            // out rows = table[idx]: 32 / (C/4) rows per warp instruction, 128-bit, line-coalesced;
            // shuffles, table reads and stores are issued in batches of 8 (memory-level parallelism)
            if (a.out != nullptr) {
                if (Cfg::TAB_SMEM && !table_ready) {   // first output phase of this warp
                    mbar_wait(bar_table, 0);
                    table_ready = true;
                }
                constexpr int LPR = C * Cfg::XB / 16;  // lanes per row (16 bytes each)
                constexpr int RPI = 32 / LPR;          // rows per iteration
                constexpr int NIT = 32 / RPI;
                constexpr int EPL = 16 / Cfg::XB;      // elements per lane
                const int sub = lane / LPR, c4 = lane % LPR;
                XT* obase = a.out + (n0 + wq * 32 + sub) * C + c4 * EPL;
                const int64_t rows_left = a.N - (n0 + wq * 32 + sub);
#pragma unroll
                for (int i0 = 0; i0 < NIT; i0 += 8) {
                    int kk[8];
                    uint4 val[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        kk[u] = __shfl_sync(0xffffffffu, bidx, (i0 + u) * RPI + sub);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if constexpr (Cfg::TAB_SMEM)
                            val[u] = *reinterpret_cast<const uint4*>(
                                smem + Cfg::OFF_TAB + (kk[u] * C + c4 * EPL) * Cfg::XB);
                        else                            // fp32 rows of the 128 KB table through L1/L2
                            val[u] = __ldg(reinterpret_cast<const uint4*>(a.table + kk[u] * C + c4 * EPL));
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if ((i0 + u) * RPI < rows_left)
                            *reinterpret_cast<uint4*>(obase + (size_t)(i0 + u) * RPI * C) = val[u];
                }
            }
            if (wq == 0) TQ_PROF(11);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sq_acc += __shfl_xor_sync(0xffffffffu, sq_acc, o);
            ties += __shfl_xor_sync(0xffffffffu, ties, o);
        }
        if (lane == 0) {
            red[warp - TQ_W_EPI] = sq_acc;             // [8]
            redt[warp - TQ_W_EPI] = ties;
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    // per-CTA partials, then the last CTA to finish reduces them in a fixed order (deterministic):
    // loss = sum / (N * 8) * commitment_cost in fp64, near-tie total in integers
    if (warp == 0) {
        uint32_t last = 0;
        if (lane == 0) {
            a.partial[blockIdx.x] =
                ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
            a.tie_partial[blockIdx.x] = ((redt[0] + redt[1]) + (redt[2] + redt[3])) +
                                        ((redt[4] + redt[5]) + (redt[6] + redt[7]));
            __threadfence();
            last = atomicAdd(a.done_counter, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            double acc = 0.0;
            uint32_t tsum = 0;
            for (int i = lane; i < (int)gridDim.x; i += 32) {
                acc += (double)__ldcg(a.partial + i);
                tsum += __ldcg(a.tie_partial + i);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
                tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
            }
            if (lane == 0) {
                *a.loss = (float)(acc * a.inv_count) * a.commitment_cost;
                if (a.near_ties != nullptr) *a.near_ties = tsum;
                *a.done_counter = 0u;                  // self-resetting: ready for the next launch
            }
        }
    }
    if (warp == TQ_W_MMA) tmem_dealloc(tmem_base, 512);
}

}  // namespace

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
                cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
}  // namespace

bool quantize_tc_supported(const vqae_quantizer_params* p, int x_layout, int out_layout,
                           bool has_out) {
    return p->w_in != nullptr && p->b_in != nullptr && p->num_codes == TQ_K && p->dim == TQ_D &&
           (p->c == 64 || p->c == 128) && x_layout == VQAE_LAYOUT_NHWC &&
           (!has_out || out_layout == VQAE_LAYOUT_NHWC);
}

size_t quantize_tc_scratch_bytes(int64_t n) {
    return (size_t)((n + TQ_M - 1) / TQ_M) * 8 + 64;
}

// Completion counters for the fused loss reduction: each launch takes the next slot of a small
// ring, so kernels of this library in flight on different streams never share one.
static unsigned int* next_done_counter() {
    constexpr int SLOTS = 64;
    static PerDevice<unsigned int*> base_dev{};       // the ring lives on the device that uses it
    static std::atomic<unsigned> next{0};
    static std::mutex mu;
    unsigned int*& base = base_dev.cur();
    if (base == nullptr) {
        std::lock_guard<std::mutex> g(mu);
        if (base == nullptr) {
            unsigned int* p = nullptr;
            if (cudaMalloc(&p, SLOTS * sizeof(unsigned int)) != cudaSuccess) return nullptr;
            if (cudaMemset(p, 0, SLOTS * sizeof(unsigned int)) != cudaSuccess) return nullptr;
            base = p;
        }
    }
    return base + (next.fetch_add(1) % SLOTS);
}

static long long* g_tq_prof = nullptr;
void quantize_tc_set_prof(long long* dev_ptr) { g_tq_prof = dev_ptr; }

template <int C, typename XT>
static int launch_quantize_tc(const vqae_quantizer_params* p, const void* x, void* out,
                              int64_t* indices, float* loss, void* scratch, uint32_t* near_ties,
                              float tie_rel_gap, float* z_out, float* diag, int64_t N, int sm_count,
                              CUtensorMapDataType tm_dtype, cudaStream_t stream) {
    using Cfg = TqCfg<C, XT>;
    if (N > 0x7fffff00ll) return VQAE_ERR_UNSUPPORTED;        // TMA row coordinate is an int32
    TqArgs<C, XT> a;
    a.x = reinterpret_cast<const XT*>(x); a.out = reinterpret_cast<XT*>(out);
    a.idx = indices; a.near_ties = near_ties;
    a.z_out = z_out; a.diag = diag; a.prof = g_tq_prof;
    a.embed = p->embed; a.table = p->table; a.N = N;
    a.num_tiles = (int)((N + TQ_M - 1) / TQ_M);
    a.tie_rel_gap = tie_rel_gap;
    a.margin = 4.f * TQ_ERR_C + 2.f * tie_rel_gap;
    a.w_in = p->w_in; a.b_in = p->b_in;
    // [N][C] viewed as a 2-D tensor; box = 128 rows x 128 bytes of channels, 128B swizzle
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return VQAE_ERR_UNSUPPORTED;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)N};
    const cuuint64_t gstride[1] = {(cuuint64_t)C * sizeof(XT)};
    const cuuint32_t box[2] = {(cuuint32_t)Cfg::CPS, (cuuint32_t)TQ_M};
    const cuuint32_t estr[2] = {1u, 1u};
    if (encode(&tmap, tm_dtype, 2, const_cast<void*>(x), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return VQAE_ERR_UNSUPPORTED;
    auto kern = quantize_tc_kernel<C, XT>;
    static PerDevice<bool> attr_set{};
    if (!attr_set.cur()) {
        VQAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)Cfg::SMEM));
        attr_set.cur() = true;
    }
    const int grid = a.num_tiles < sm_count ? a.num_tiles : sm_count;
    a.partial = reinterpret_cast<float*>(scratch);
    a.tie_partial = reinterpret_cast<uint32_t*>(scratch) + grid;
    a.done_counter = next_done_counter();
    if (a.done_counter == nullptr) return VQAE_ERR_CUDA;
    a.loss = loss;
    a.inv_count = 1.0 / ((double)N * (double)TQ_D);
    a.commitment_cost = p->commitment_cost;
    kern<<<grid, TQ_THREADS, Cfg::SMEM, stream>>>(a, tmap);
    return check_launch();
}

// x / out: NHWC, both of element type `dtype` (VQAE_DT_F32 / VQAE_DT_BF16 / VQAE_DT_F16)
int quantize_tc(const vqae_quantizer_params* p, const void* x, void* out, int dtype,
                int64_t* indices, float* loss, void* scratch, uint32_t* near_ties,
                float tie_rel_gap, float* z_out, float* diag, int64_t N, int sm_count,
                cudaStream_t stream) {
#define TQ_LAUNCH(CC, T, TM)                                                                      \
    return launch_quantize_tc<CC, T>(p, x, out, indices, loss, scratch, near_ties, tie_rel_gap,   \
                                     z_out, diag, N, sm_count, TM, stream)
    if (p->c == 64) {
        if (dtype == VQAE_DT_F32) TQ_LAUNCH(64, float, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        if (dtype == VQAE_DT_BF16) TQ_LAUNCH(64, __nv_bfloat16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
        if (dtype == VQAE_DT_F16) TQ_LAUNCH(64, __half, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    } else if (p->c == 128) {
        if (dtype == VQAE_DT_F32) TQ_LAUNCH(128, float, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        if (dtype == VQAE_DT_BF16) TQ_LAUNCH(128, __nv_bfloat16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
        if (dtype == VQAE_DT_F16) TQ_LAUNCH(128, __half, CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    }
#undef TQ_LAUNCH
    return VQAE_ERR_UNSUPPORTED;
}

int quantize_tc_f32(const vqae_quantizer_params* p, const float* x, float* out, int64_t* indices,
                    float* loss, void* scratch, uint32_t* near_ties, float tie_rel_gap,
                    float* z_out, float* diag, int64_t N, int sm_count, cudaStream_t stream) {
    return quantize_tc(p, x, out, VQAE_DT_F32, indices, loss, scratch, near_ties, tie_rel_gap, z_out,
                       diag, N, sm_count, stream);
}

}  // namespace vqae
