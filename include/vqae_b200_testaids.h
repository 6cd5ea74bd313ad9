/* vqae_b200_testaids.h -- C ABI of libvqae_b200_testaids.so: test and measurement aids of the
 * sm_100a VQ-AE kernels.  NOT part of the product boundary (include/vqae_b200.h); used by tests/
 * and profiles/ only.  Links against libvqae_b200.so.
 */
#ifndef VQAE_B200_TESTAIDS_H_
#define VQAE_B200_TESTAIDS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* profiling aid: DEVICE int64 [8][32] that CTA 0 fills with clock64() stamps of eight steady-state
 * half-rounds (MMA warp: slots 0-6, worker warp 0: slots 8-22); NULL switches it off          */
void vqae_trunk_resident_set_profile(long long* phase_clocks);

/* same call, additionally writing clock64() at the 8 phase boundaries of every CTA's first tile
 * to phase_clocks[grid][8] (device memory, >= 8 * 4 * SM-count int64) -- profiling aid        */
int vqae_same_block_f16_profile(const void* x, void* out, int io_dtype, const void* w_packed,
                                 const float* scalars8_host, int64_t batch, int height, int width,
                                 int c, long long* phase_clocks, void* stream);
/* tcgen05.mma issue-rate microbenchmark (timing aid): `reps` MMAs of 128 x n x 16 bf16 from shared
 * memory in layout_type 0 (un-swizzled K-major) or 2 (128-byte swizzle); out2[0] = cycles, [1] = reps */
int vqae_tc_mma_bench(int n, int layout_type, int reps, int a_stride_rows, long long* out2,
                      void* stream);
/* the same measurement with n_issuers warps per CTA issuing concurrently (own accumulators) and
 * ctas_per_sm CTAs resident per SM; m in {64, 128}; mode bit 0: every issuer has its own A and B
 * regions, bit 1: A row groups 160 B apart (the resident kernel's tile layout); out_per_cta: DEVICE int64 [SMs * ctas_per_sm]
 * cycles until every issuer's `reps` MMAs have completed                                       */
int vqae_tc_mma_bench2(int m, int n, int reps, int n_issuers, int ctas_per_sm, int mode,
                       long long* out_per_cta, void* stream);
/* descriptor/TMEM self test: d[128][64] = a[row_shift + m][0..63] . b[n][0..63] (bf16 in, fp32 out),
 * a: [a_rows][64] bf16 row-major, b: [64][64] bf16 row-major (device pointers)                */
int vqae_tc_selftest(const void* a_bf16, int a_rows, int row_shift, const void* b_bf16, float* d,
                     void* stream);

/* profiling aid: while non-NULL, CTA 0 of every tcgen05 quantiser launch writes clock64() stamps of
 * its tiles 10..13 to phase_clocks[4][16] (device memory, 64 int64); NULL switches it off      */
void vqae_quantize_tc_set_profile(long long* phase_clocks);

#ifdef __cplusplus
}
#endif
#endif /* VQAE_B200_TESTAIDS_H_ */
