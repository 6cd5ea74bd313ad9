// Inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path: mbarrier, bulk async copy
// (TMA without tensor map), tcgen05 alloc / mma / commit / ld, and the shared-memory matrix
// descriptor for the un-swizzled K-major canonical layout.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

// GEMM operand element type of the reduced-precision conv kernels: IEEE fp16 (11-bit significand).
// The reference's own inference runs under torch.autocast('cuda') = fp16
// (scripts/extract_embeddings/extract_embeddings.py:124), and with fp16 operands the encoder flips
// ~5x fewer codes against the reference's fp32 run than with bf16 operands (8-bit significand) at the
// same tensor-core rate.  0 selects bf16 (wider exponent range) for the whole library.
#ifndef VQAE_OPERAND_F16
#define VQAE_OPERAND_F16 1
#endif

namespace vqae {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
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
    } while (!done);
}

// ---- proxies / fences ---------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- bulk async copy global -> shared (UBLKCP), completion on an mbarrier -------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(dst_smem),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// ---- tensor memory ----------------------------------------------------------------------------
// one full warp; writes the allocated base address (lane 0 / column base) to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}

// 32 lanes x 32 columns of fp32: thread t of the warp receives lane (lane_base + t), 32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]) {
    static_assert(N == 16 || N == 32, "tmem_ld: 16 or 32 columns");
    if constexpr (N == 32) tmem_ld32(taddr, v);
    else tmem_ld16(taddr, v);
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- UMMA -------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"):
//   element (row r, k) of the operand lives at
//     start + (r % 8) * 16 + (r / 8) * SBO + (k / 8) * LBO + (k % 8) * 2        [bf16]
// i.e. 8x8 core matrices of 128 contiguous bytes; SBO steps 8 rows, LBO steps 8 k-elements.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;  // descriptor version (Blackwell)
    return d;         // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// instruction descriptor: kind::f16, A = B = bf16 (K-major), D = fp32, shape M x N x 16
// (a_format / b_format: 0 = F16, 1 = BF16)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (VQAE_OPERAND_F16 ? 0u : ((1u << 7) | (1u << 10))) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// always bf16 operands (the quantiser's split-bf16 filter GEMM, the self test)
__host__ __device__ constexpr uint32_t make_idesc_true_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T.  Called by ALL lanes of the (converged) MMA warp with
// warp-uniform operands, so that descriptor arithmetic stays in uniform registers; only the lane
// with `leader` set issues the instruction.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate, uint32_t leader) {
    (void)leader;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of the elected thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar, uint32_t leader) {
    (void)leader;
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(bar)
        : "memory");
}

// two fp32 -> two bf16, whatever the conv operand type (the quantiser's split-bf16 filter GEMM)
__device__ __forceinline__ uint32_t pack_true_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// two fp32 -> two operand elements (fp16, or bf16 when VQAE_OPERAND_F16 == 0)
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
#if VQAE_OPERAND_F16
    __half2 v = __floats2half2_rn(lo, hi);
#else
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
#endif
    return *reinterpret_cast<uint32_t*>(&v);
}

// ELU(alpha=1) for the bf16 path: x > 0 ? x : e^x - 1 with the fast exponential
// (one FMUL + one MUFU.EX2, flush-to-zero: no denormal fix-up code around the SFU call)
__device__ __forceinline__ float elu_fast(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return x > 0.f ? x : e - 1.f;
}

// 8 fp32 -> 8 operand elements of  elu(v + pre) + post  (the Fixup pre-activation, conv_block.py:199-208)
// with packed fp32x2 arithmetic: elu(v + pre) + post = (v > -pre) ? v + (pre + post)
//                                                                  : exp2(v * log2e + pre * log2e) + (post - 1)
// -- per pair FADD2, FFMA2, 2 x MUFU.EX2, FADD2, 2 x FSETP / FSEL, F2FP = 5 issue slots per element
// against 9 for the element-wise form (the worker units of the resident kernel are issue bound:
// DESIGN.md section 4.5).  Same formula as mma::ActC (mma_common.cuh).
__device__ __forceinline__ uint32_t act_pair(float x0, float x1, float2 sum, float2 tl, float2 pm1,
                                             float npre) {
    const float2 x = make_float2(x0, x1);
    const float2 lin = __fadd2_rn(x, sum);
    const float2 tt = __ffma2_rn(x, make_float2(1.4426950408889634f, 1.4426950408889634f), tl);
    float2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(tt.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(tt.y));
    const float2 ex = __fadd2_rn(e, pm1);
    return pack_bf16(x0 > npre ? lin.x : ex.x, x1 > npre ? lin.y : ex.y);
}
__device__ __forceinline__ uint4 act_pack8(const float* v, float pre, float post) {
    const float s = pre + post, t = pre * 1.4426950408889634f, m = post - 1.f, npre = -pre;
    const float2 sum = make_float2(s, s), tl = make_float2(t, t), pm1 = make_float2(m, m);
    uint4 o;
    o.x = act_pair(v[0], v[1], sum, tl, pm1, npre);
    o.y = act_pair(v[2], v[3], sum, tl, pm1, npre);
    o.z = act_pair(v[4], v[5], sum, tl, pm1, npre);
    o.w = act_pair(v[6], v[7], sum, tl, pm1, npre);
    return o;
}
// 8 fp32 -> 8 bf16 of  v + add  (skip path: no activation, conv_block.py:211-213)
__device__ __forceinline__ uint4 add_pack8(const float* v, float add) {
    uint4 o;
    o.x = pack_bf16(v[0] + add, v[1] + add);
    o.y = pack_bf16(v[2] + add, v[3] + add);
    o.z = pack_bf16(v[4] + add, v[5] + add);
    o.w = pack_bf16(v[6] + add, v[7] + add);
    return o;
}

// ---- activation stream I/O: the NHWC tensors between kernels are fp32 or fp16 -----------------
// (fp16 in the reduced-precision path: half the HBM bytes of the block-boundary tensors; the
// arithmetic between load and store is fp32 either way)
template <typename T>
struct StreamIO;
template <>
struct StreamIO<float> {
    static constexpr int DT = 0;                          // VQAE_DT_F32
    __device__ static __forceinline__ void load8(const float* p, float (&v)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ static __forceinline__ float4 load4(const float* p) {
        return __ldg(reinterpret_cast<const float4*>(p));
    }
    __device__ static __forceinline__ void store4(float* p, float4 v) {
        *reinterpret_cast<float4*>(p) = v;
    }
};
template <>
struct StreamIO<__half> {
    static constexpr int DT = 3;                          // VQAE_DT_F16
    __device__ static __forceinline__ void load8(const __half* p, float (&v)[8]) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    __device__ static __forceinline__ float4 load4(const __half* p) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    __device__ static __forceinline__ void store4(__half* p, float4 v) {
        const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t*>(&a);
        u.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = u;
    }
};

}  // namespace tc
}  // namespace vqae
