// Shared pieces of the warp-level-MMA kernels (mma_same.cu, mma_down.cu): mma.sync / ldmatrix
// wrappers, the packed-fp32 ELU, multiplication-based integer division.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "tc_common.cuh"

namespace vqae {
namespace mma {

// n / d for 0 <= n < 2^31 by a multiplication: q = umulhi(n, ceil(2^32 / d)), at most one too large
struct FastDiv {
    uint32_t d, m;
    __device__ __forceinline__ int div(int n) const {
        uint32_t q = __umulhi((uint32_t)n, m);
        q -= (q * d > (uint32_t)n) ? 1u : 0u;
        return (int)q;
    }
};
inline FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = (uint32_t)d;
    f.m = d <= 1 ? 0xffffffffu : (uint32_t)((0x100000000ull + (uint64_t)d - 1) / (uint64_t)d);
    return f;
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// f16x2( elu(v + pre) + post ) for a pair, with packed fp32x2 arithmetic: six issue slots per pair
// besides the two MUFU.EX2 (FADD2, FFMA2, FADD2, 2 x FSETP/FSEL fused by ptxas into predicated
// selects).  elu(v + pre) + post = (v > -pre) ? v + (pre + post) : exp(v + pre) + (post - 1).
struct ActC {
    float2 sum, l2e, tl, pm1;
    float npre;
    __device__ __forceinline__ ActC(float pre, float post) {
        constexpr float L2E = 1.4426950408889634f;
        sum = make_float2(pre + post, pre + post);
        l2e = make_float2(L2E, L2E);
        tl = make_float2(pre * L2E, pre * L2E);
        pm1 = make_float2(post - 1.f, post - 1.f);
        npre = -pre;
    }
    __device__ __forceinline__ uint32_t operator()(float x0, float x1) const {
        const float2 x = make_float2(x0, x1);
        const float2 lin = __fadd2_rn(x, sum);
        const float2 tt = __ffma2_rn(x, l2e, tl);
        float2 e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(tt.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(tt.y));
        const float2 ex = __fadd2_rn(e, pm1);
        return pack_h2(x0 > npre ? lin.x : ex.x, x1 > npre ? lin.y : ex.y);
    }
};

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

}  // namespace mma
}  // namespace vqae
