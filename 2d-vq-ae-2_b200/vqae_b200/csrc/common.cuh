// Shared helpers for the sm_100a kernels of the VQ-AE inference path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vqae_b200.h"

namespace vqae {

// thread-local record of the last CUDA error (vqae_last_cuda_error)
void set_last_cuda_error(cudaError_t e);
void count_launch(int n = 1);

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_cuda_error(e);
        return VQAE_ERR_CUDA;
    }
    count_launch();
    return VQAE_OK;
}

#define VQAE_CUDA_TRY(expr)                         \
    do {                                            \
        cudaError_t _e = (expr);                    \
        if (_e != cudaSuccess) {                    \
            ::vqae::set_last_cuda_error(_e);        \
            return VQAE_ERR_CUDA;                   \
        }                                           \
    } while (0)

// nn.ELU(alpha=1) (conf/model/layers/activation/elu.yaml): x > 0 ? x : expm1(x)
__device__ __forceinline__ float elu1(float x) { return x > 0.0f ? x : expm1f(x); }

// the pre-activation in front of every branch conv of PreActFixupResBlock
// (layers/conv_block.py:199-208):  conv(act(x + a) + b);  act optional (skip path has none)
struct PreOp {
    float a;
    float b;
    int use_elu;
    __device__ __forceinline__ float operator()(float x) const {
        float v = x + a;
        if (use_elu) v = elu1(v);
        return v + b;
    }
};

// Per-device caches: function attributes, occupancy results and scratch allocations belong to ONE
// device; a process that drives several GPUs must not reuse them across devices.
constexpr int kMaxDevices = 64;
inline int current_device_index() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) return 0;
    return d;
}
template <typename T>
struct PerDevice {
    T v[kMaxDevices];
    T& cur() { return v[current_device_index()]; }
};

inline unsigned ceil_div_u(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace vqae
