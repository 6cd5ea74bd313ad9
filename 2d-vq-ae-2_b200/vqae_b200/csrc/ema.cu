// Training-mode codebook maintenance of EMAVectorQuantizer (layers/vq.py:47-94; scope row f-4) and the
// elementwise sum the multi-level hierarchy needs (model.py:208, 283-288):
//
//   _update_ema (vq.py:47-74)
//       new_cluster_size[k] = #{n : idx[n] = k}                     one_hot.sum(0)
//       dw[k, :]            = sum_{n : idx[n] = k} z[n, :]          one_hot.T @ flat_input
//       [all-reduce of both across ranks: done by the host between the two entry points]
//       cluster_size = decay * cluster_size + (1 - decay) * new_cluster_size
//       embed_avg    = decay * embed_avg    + (1 - decay) * dw
//       n = sum(cluster_size);  smoothed = n * (cluster_size + alpha) / (n + K alpha)
//       embed = embed_avg / smoothed[:, None]
//   _init_ema (vq.py:76-94)
//       mean / unbiased std of z over the rows, embed = embed * std + mean, embed_avg = embed,
//       cluster_size += N / K
//
// The reference materialises an N x K one-hot matrix (537 MB at N = 524 288) and runs a GEMM over it.
// Here nothing of size N x K exists: a CTA stages a chunk of rows (indices + an 8-column slice of z) in
// shared memory and every thread owns ONE code, scanning the chunk in row order -- no atomics, a fixed
// summation order, bit-identical from run to run.  Chunk partials are summed in fp64 in chunk order.
// HBM-bound byte work: z and the indices are read once (N (4 D + 8) bytes).
#include "common.cuh"
#include "kernels.cuh"

namespace vqae {
namespace {

constexpr int EMA_ROWS = 1024;      // rows per chunk
constexpr int EMA_THREADS = 256;
constexpr int EMA_DG = 8;           // z columns per CTA (blockIdx.y selects the group)

// partial_dw [n_chunks][K][D] fp32, partial_cnt [n_chunks][K] int32
__global__ void __launch_bounds__(EMA_THREADS)
ema_accumulate_kernel(const float* __restrict__ z, const int64_t* __restrict__ idx, int64_t n, int K,
                      int D, float* __restrict__ partial_dw, int* __restrict__ partial_cnt) {
    __shared__ int s_idx[EMA_ROWS];
    __shared__ __align__(16) float s_z[EMA_ROWS][EMA_DG];
    const int chunk = blockIdx.x, d0 = blockIdx.y * EMA_DG;
    const int64_t r0 = (int64_t)chunk * EMA_ROWS;
    const int rows = (int)min((int64_t)EMA_ROWS, n - r0);
    const int dn = min(EMA_DG, D - d0);
    for (int r = threadIdx.x; r < rows; r += EMA_THREADS) s_idx[r] = (int)__ldg(idx + r0 + r);
    for (int e = threadIdx.x; e < rows * EMA_DG; e += EMA_THREADS) {
        const int r = e / EMA_DG, d = e % EMA_DG;
        s_z[r][d] = d < dn ? __ldg(z + (r0 + r) * D + d0 + d) : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += EMA_THREADS) {
        float acc[EMA_DG];
#pragma unroll
        for (int d = 0; d < EMA_DG; ++d) acc[d] = 0.f;
        int cnt = 0;
        for (int r = 0; r < rows; ++r) {
            if (s_idx[r] == k) {                      // rows in order: a fixed summation order
                const float4 a = *reinterpret_cast<const float4*>(&s_z[r][0]);
                const float4 b = *reinterpret_cast<const float4*>(&s_z[r][4]);
                acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
                acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
                ++cnt;
            }
        }
        float* o = partial_dw + ((size_t)chunk * K + k) * D + d0;
        for (int d = 0; d < dn; ++d) o[d] = acc[d];
        if (blockIdx.y == 0) partial_cnt[(size_t)chunk * K + k] = cnt;
    }
}

// one thread per (k, d) with d = D standing for the count column
__global__ void ema_reduce_kernel(const float* __restrict__ partial_dw, const int* __restrict__ partial_cnt,
                                  int n_chunks, int K, int D, float* __restrict__ counts,
                                  float* __restrict__ dw) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < K * D) {
        double s = 0.0;
        for (int c = 0; c < n_chunks; ++c) s += (double)partial_dw[(size_t)c * K * D + t];
        dw[t] = (float)s;
    } else if (t < K * D + K) {
        const int k = t - K * D;
        long long s = 0;
        for (int c = 0; c < n_chunks; ++c) s += partial_cnt[(size_t)c * K + k];
        counts[k] = (float)s;                           // the reference's one-hot sum is fp32 too
    }
}

// single CTA: the buffers are K x D with K, D of a few hundred at most
__global__ void __launch_bounds__(256)
ema_update_kernel(float* __restrict__ embed, float* __restrict__ embed_avg, float* __restrict__ cluster_size,
                  const float* __restrict__ counts, const float* __restrict__ dw, int K, int D, float decay,
                  float one_minus_decay, float alpha, float k_alpha) {
    __shared__ double s_part[256];
    __shared__ float s_n;
    double part = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        // cluster_size.mul_(decay).add_(new, alpha = 1 - decay)   (vq.py:60-62)
        const float cs = fmaf(counts[k], one_minus_decay, cluster_size[k] * decay);
        cluster_size[k] = cs;
        part += (double)cs;
    }
    s_part[threadIdx.x] = part;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) s_part[threadIdx.x] += s_part[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) s_n = (float)s_part[0];
    __syncthreads();
    const float n = s_n;
    const float denom = n + k_alpha;                                  // vq.py:70
    for (int e = threadIdx.x; e < K * D; e += blockDim.x) {
        const int k = e / D;
        const float avg = fmaf(dw[e], one_minus_decay, embed_avg[e] * decay);   // vq.py:64
        embed_avg[e] = avg;
        const float smoothed = n * ((cluster_size[k] + alpha) / denom);         // vq.py:67-71
        embed[e] = avg / smoothed;                                              // vq.py:73-74
    }
}

// column sums for mean / unbiased std: partial [n_chunks][2][D] fp64
constexpr int STAT_ROWS = 4096;
__global__ void __launch_bounds__(256)
column_stats_partial_kernel(const float* __restrict__ z, int64_t n, int D, int DP, double* __restrict__ partial) {
    __shared__ double s_sum[256], s_sq[256];
    const int d = threadIdx.x % DP, lane = threadIdx.x / DP, lanes = 256 / DP;
    const int64_t r0 = (int64_t)blockIdx.x * STAT_ROWS;
    const int rows = (int)min((int64_t)STAT_ROWS, n - r0);
    for (int dbase = 0; dbase < D; dbase += DP) {
        double s = 0.0, q = 0.0;
        if (dbase + d < D)
            for (int r = lane; r < rows; r += lanes) {
                const double v = (double)__ldg(z + (r0 + r) * D + dbase + d);
                s += v;
                q += v * v;
            }
        s_sum[threadIdx.x] = s;
        s_sq[threadIdx.x] = q;
        __syncthreads();
        if (lane == 0 && dbase + d < D) {
            for (int l = 1; l < lanes; ++l) { s += s_sum[l * DP + d]; q += s_sq[l * DP + d]; }
            partial[((size_t)blockIdx.x * 2 + 0) * D + dbase + d] = s;
            partial[((size_t)blockIdx.x * 2 + 1) * D + dbase + d] = q;
        }
        __syncthreads();
    }
}
__global__ void column_stats_final_kernel(const double* __restrict__ partial, int n_chunks, int64_t n, int D,
                                          float* __restrict__ mean, float* __restrict__ stdv) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    double s = 0.0, q = 0.0;
    for (int c = 0; c < n_chunks; ++c) {
        s += partial[((size_t)c * 2 + 0) * D + d];
        q += partial[((size_t)c * 2 + 1) * D + d];
    }
    const double m = s / (double)n;
    const double var = (q - (double)n * m * m) / (double)(n - 1);       // torch.std: correction = 1
    mean[d] = (float)m;
    stdv[d] = (float)sqrt(var > 0.0 ? var : 0.0);
}

__global__ void ema_init_kernel(float* __restrict__ embed, float* __restrict__ embed_avg,
                                float* __restrict__ cluster_size, const float* __restrict__ mean,
                                const float* __restrict__ stdv, int K, int D, float cluster_add) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < K * D) {
        const int d = e % D;
        const float v = embed[e] * stdv[d] + mean[d];     // embed.mul_(std); embed.add_(mean)  (vq.py:90-91)
        embed[e] = v;
        embed_avg[e] = v;                                 // vq.py:92
    }
    if (e < K) cluster_size[e] += cluster_add;            // vq.py:94
}

__global__ void add_f32_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out,
                               int64_t n4, const float* a1, const float* b1, float* o1, int tail) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 x = __ldg(a + i), y = __ldg(b + i);
        out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) o1[threadIdx.x] = a1[threadIdx.x] + b1[threadIdx.x];
}

inline int ema_chunks(int64_t n) { return (int)((n + EMA_ROWS - 1) / EMA_ROWS); }
inline int stat_chunks(int64_t n) { return (int)((n + STAT_ROWS - 1) / STAT_ROWS); }
inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace
}  // namespace vqae

using namespace vqae;

extern "C" {

size_t vqae_ema_scratch_bytes(int64_t n, int num_codes, int dim) {
    if (n <= 0 || num_codes <= 0 || dim <= 0) return 0;
    const size_t acc = al256((size_t)ema_chunks(n) * num_codes * dim * sizeof(float)) +
                       al256((size_t)ema_chunks(n) * num_codes * sizeof(int));
    const size_t stat = al256((size_t)stat_chunks(n) * 2 * dim * sizeof(double));
    return acc > stat ? acc : stat;
}

int vqae_ema_accumulate_f32(const float* z, const int64_t* indices, int64_t n, int num_codes, int dim,
                            float* counts, float* dw, void* scratch, size_t scratch_bytes, void* stream) {
    if (!z || !indices || !counts || !dw || n <= 0 || num_codes <= 0 || dim <= 0) return VQAE_ERR_BAD_ARG;
    if (!scratch || scratch_bytes < vqae_ema_scratch_bytes(n, num_codes, dim)) return VQAE_ERR_SCRATCH;
    const int nc = ema_chunks(n);
    float* pdw = reinterpret_cast<float*>(scratch);
    int* pcnt = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(scratch) +
                                       al256((size_t)nc * num_codes * dim * sizeof(float)));
    cudaStream_t st = (cudaStream_t)stream;
    ema_accumulate_kernel<<<dim3(nc, (dim + EMA_DG - 1) / EMA_DG), EMA_THREADS, 0, st>>>(
        z, indices, n, num_codes, dim, pdw, pcnt);
    if (int rc = check_launch()) return rc;
    const int total = num_codes * dim + num_codes;
    ema_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(pdw, pcnt, nc, num_codes, dim, counts, dw);
    return check_launch();
}

int vqae_ema_update_f32(float* embed, float* embed_avg, float* cluster_size, const float* counts,
                        const float* dw, int num_codes, int dim, float decay, float laplace_alpha,
                        void* stream) {
    if (!embed || !embed_avg || !cluster_size || !counts || !dw || num_codes <= 0 || dim <= 0)
        return VQAE_ERR_BAD_ARG;
    // the host-side scalars are formed in double like the reference's Python floats (vq.py:61-70)
    const float omd = (float)(1.0 - (double)decay);
    const float k_alpha = (float)((double)num_codes * (double)laplace_alpha);
    ema_update_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(embed, embed_avg, cluster_size, counts, dw,
                                                           num_codes, dim, decay, omd, laplace_alpha, k_alpha);
    return check_launch();
}

int vqae_column_stats_f32(const float* z, int64_t n, int dim, float* mean, float* std_unbiased,
                          void* scratch, size_t scratch_bytes, void* stream) {
    if (!z || !mean || !std_unbiased || n < 2 || dim <= 0) return VQAE_ERR_BAD_ARG;
    if (!scratch || scratch_bytes < al256((size_t)stat_chunks(n) * 2 * dim * sizeof(double)))
        return VQAE_ERR_SCRATCH;
    int dp = 1;
    while (dp < dim && dp < 256) dp *= 2;
    const int nc = stat_chunks(n);
    double* partial = reinterpret_cast<double*>(scratch);
    cudaStream_t st = (cudaStream_t)stream;
    column_stats_partial_kernel<<<nc, 256, 0, st>>>(z, n, dim, dp, partial);
    if (int rc = check_launch()) return rc;
    column_stats_final_kernel<<<(dim + 127) / 128, 128, 0, st>>>(partial, nc, n, dim, mean, std_unbiased);
    return check_launch();
}

int vqae_ema_init_f32(float* embed, float* embed_avg, float* cluster_size, const float* mean,
                      const float* std_unbiased, int num_codes, int dim, float cluster_add, void* stream) {
    if (!embed || !embed_avg || !cluster_size || !mean || !std_unbiased || num_codes <= 0 || dim <= 0)
        return VQAE_ERR_BAD_ARG;
    const int total = num_codes * dim;
    ema_init_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        embed, embed_avg, cluster_size, mean, std_unbiased, num_codes, dim, cluster_add);
    return check_launch();
}

int vqae_add_f32(const float* a, const float* b, float* out, int64_t n, void* stream) {
    if (!a || !b || !out || n < 0) return VQAE_ERR_BAD_ARG;
    if (n == 0) return VQAE_OK;
    if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) return VQAE_ERR_BAD_ARG;
    const int64_t n4 = n / 4;
    const int tail = (int)(n - n4 * 4);
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    add_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
        reinterpret_cast<float4*>(out), n4, a + n4 * 4, b + n4 * 4, out + n4 * 4, tail);
    return check_launch();
}

}  // extern "C"
