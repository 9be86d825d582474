// common.cuh -- shared helpers for libb200cc (error reporting, launch accounting, reductions)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/b200cc.h"

namespace b200cc {

typedef long long i64;

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

#define B200CC_CUDA_OK(expr)                                                     \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      b200cc::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));         \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

int sm_count();

// ---- block-wide sum of a double; result valid in thread 0 --------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sm /* >= NT/32 doubles */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (lane < NT / 32) ? sm[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// second stage of every deterministic reduction: out[q] (=|+=) scale * sum_p partial[q*stride + p]
__global__ void final_reduce_kernel(const double* partial, int nparts, int stride, double* out,
                                    int accumulate, double scale);
int launch_final_reduce(const double* partial, int nparts, int stride, int nout, double* out,
                        int accumulate, double scale, cudaStream_t st);

}  // namespace b200cc
