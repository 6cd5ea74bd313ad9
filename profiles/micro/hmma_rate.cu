// Legacy warp-level mma.sync.m16n8k16 (fp16 in, fp32 accumulate) issue rate on sm_100a:
// `warps` warps per CTA, one CTA per SM, ILP independent accumulator chains per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int ILP>
__global__ void k(int reps, float* out, long long* cyc) {
    uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
    float c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = (float)threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int reps = 4000;
    for (int warps : {4, 8, 16, 32}) {
        k<8><<<148, warps * 32>>>(reps, out, cyc);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double c = (double)h[0];
        double macs = (double)reps * 8 * warps * 16 * 8 * 16;
        printf("warps %2d: %.0f cycles, %.1f MAC/clk/SM (%.2f cycles per HMMA per SMSP), err %s\n", warps, c, macs / c,
               c / (reps * 8.0 * warps / 4.0), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
