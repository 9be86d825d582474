// elementwise.cu -- the HBM-bound passes of the CCSD iteration, fused and deterministic.
//   build_tau, denominators, the fused {symmetrise r2, Jacobi update, rms} pass, the energy dot,
//   DIIS multi-dot / multi-axpy.  All reductions are two-stage (per-CTA partials in fixed order,
//   then one CTA sums them): bitwise reproducible, no atomics.
#include <stdarg.h>
#include "common.cuh"

namespace b200cc {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

__global__ void final_reduce_kernel(const double* partial, int nparts, int stride, double* out,
                                    int accumulate, double scale) {
  __shared__ double sm[8];
  const double* p = partial + (i64)blockIdx.x * stride;
  double s = 0.0;
  for (int e = threadIdx.x; e < nparts; e += 256) s += p[e];
  s = block_sum<256>(s, sm);
  if (threadIdx.x == 0) out[blockIdx.x] = accumulate ? out[blockIdx.x] + scale * s : scale * s;
}

int launch_final_reduce(const double* partial, int nparts, int stride, int nout, double* out,
                        int accumulate, double scale, cudaStream_t st) {
  final_reduce_kernel<<<nout, 256, 0, st>>>(partial, nparts, stride, out, accumulate, scale);
  return check_launch("final_reduce_kernel");
}

// ---- tau = f1*t2 + f2*t1 (x) t1 ------------------------------------------------------------------
__global__ void __launch_bounds__(256) tau_kernel(int no, int nv, double f1, double f2,
                                                  const double* __restrict__ t1, const double* __restrict__ t2,
                                                  double* __restrict__ tau) {
  const int vv = nv * nv;
  for (int ij = blockIdx.x; ij < no * no; ij += gridDim.x) {
    const int i = ij / no, j = ij - i * no;
    const double* ti = t1 + (i64)i * nv;
    const double* tj = t1 + (i64)j * nv;
    const i64 base = (i64)ij * vv;
    for (int ab = threadIdx.x; ab < vv; ab += blockDim.x) {
      const int a = ab / nv, b = ab - a * nv;
      tau[base + ab] = f1 * t2[base + ab] + f2 * ti[a] * tj[b];
    }
  }
}

__global__ void __launch_bounds__(256) div_d2_kernel(int no, int nv, const double* __restrict__ eo,
                                                     const double* __restrict__ ev, const double* in, double* out) {
  const int vv = nv * nv;
  for (int ij = blockIdx.x; ij < no * no; ij += gridDim.x) {
    const int i = ij / no, j = ij - i * no;
    const double eij = eo[i] + eo[j];
    const i64 base = (i64)ij * vv;
    for (int ab = threadIdx.x; ab < vv; ab += blockDim.x) {
      const int a = ab / nv, b = ab - a * nv;
      out[base + ab] = in[base + ab] / (eij - ev[a] - ev[b]);
    }
  }
}

__global__ void __launch_bounds__(256) div_d1_kernel(int no, int nv, const double* __restrict__ eo,
                                                     const double* __restrict__ ev, const double* in, double* out) {
  const int n = no * nv;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int i = e / nv, a = e - i * nv;
    out[e] = in[e] / (eo[i] - ev[a]);
  }
}

// ---- fused: r2 = half + half^T ; t += r/D ; sum (r/D)^2 -------------------------------------------
// Work item = (i<=j, tile A, tile B) of 32x32 virtual tiles together with its partner (j,i,B,A).
// block = 32 x 8 threads.  Partial sums -> scratch[blockIdx.x].
template <bool WRITE_R2>
__global__ void __launch_bounds__(256) sym_update_kernel(int no, int nv, const double* __restrict__ eo,
                                                         const double* __restrict__ ev, const double* __restrict__ r1,
                                                         double* r2, double* t1, double* t2, double* scratch,
                                                         int update) {
  __shared__ double X[32][33], Y[32][33], R[32][33];
  __shared__ double red[8];
  const int nt = (nv + 31) / 32;
  const i64 npair = (i64)no * (no + 1) / 2;
  const i64 nwork = npair * nt * nt;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const i64 vv = (i64)nv * nv;
  double part = 0.0;
  for (i64 w = blockIdx.x; w < nwork; w += gridDim.x) {
    i64 r = w;
    const int tb = (int)(r % nt); r /= nt;
    const int ta = (int)(r % nt); r /= nt;
    // r -> (i <= j): row-major over the upper triangle
    int i = 0;
    {
      i64 rem = r;
      // rows have lengths no, no-1, ...
      while (rem >= no - i) { rem -= no - i; ++i; }
      r = rem;
    }
    const int j = i + (int)r;
    const bool diag_ij = (i == j);
    if (diag_ij && ta > tb) continue;  // covered by the (tb, ta) item (uniform branch)
    const bool self = diag_ij && ta == tb;
    const int a0 = ta * 32, b0 = tb * 32;
    const double* Hij = r2 + ((i64)i * no + j) * vv;
    const double* Hji = r2 + ((i64)j * no + i) * vv;
    __syncthreads();  // previous iteration's shared tiles fully consumed
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const int y = ty + k;
      // X[y][x] = H_ij[a0+y][b0+x]
      X[y][tx] = (a0 + y < nv && b0 + tx < nv) ? Hij[(i64)(a0 + y) * nv + b0 + tx] : 0.0;
      // Y[y][x] = H_ji[b0+y][a0+x]
      Y[y][tx] = (b0 + y < nv && a0 + tx < nv) ? Hji[(i64)(b0 + y) * nv + a0 + tx] : 0.0;
    }
    __syncthreads();
    const double eij = update ? eo[i] + eo[j] : 0.0;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const int y = ty + k;
      const double rv = X[y][tx] + Y[tx][y];  // r_ij[a0+y][b0+x]
      R[y][tx] = rv;
    }
    __syncthreads();
    // element (i,j,a0+y,b0+x)
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const int y = ty + k;
      const int a = a0 + y, b = b0 + tx;
      if (a < nv && b < nv) {
        const double rv = R[y][tx];
        const i64 off = ((i64)i * no + j) * vv + (i64)a * nv + b;
        if (WRITE_R2) r2[off] = rv;
        if (update) {
          const double d = rv / (eij - ev[a] - ev[b]);
          t2[off] += d;
          part += self ? d * d : 2.0 * d * d;
        }
      }
    }
    if (!self) {
      // partner element (j,i,b0+y,a0+x) carries r_ij[a0+x][b0+y] = R[x][y]
#pragma unroll
      for (int k = 0; k < 32; k += 8) {
        const int y = ty + k;
        const int b = b0 + y, a = a0 + tx;
        if (a < nv && b < nv) {
          const double rv = R[tx][y];
          const i64 off = ((i64)j * no + i) * vv + (i64)b * nv + a;
          if (WRITE_R2) r2[off] = rv;
          if (update) {
            const double d = rv / (eij - ev[a] - ev[b]);
            t2[off] += d;
          }
        }
      }
    }
  }
  // singles (tiny): CTA 0
  if (blockIdx.x == 0 && r1 != nullptr) {
    const int tid = ty * 32 + tx;
    for (int e = tid; e < no * nv; e += 256) {
      const int i = e / nv, a = e - i * nv;
      const double d = r1[e] / (eo[i] - ev[a]);
      if (update) t1[e] += d;
      part += d * d;
    }
  }
  __syncthreads();
  // block reduce (threads are 2-D: linearise)
  {
    const int tid = ty * 32 + tx;
    double v = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
      double s = tid < 8 ? red[tid] : 0.0;
      s = warp_sum(s);
      if (tid == 0) scratch[blockIdx.x] = s;
    }
  }
}

// r2 already symmetric: t2 += r2/D, sum d^2
__global__ void __launch_bounds__(256) update_plain_kernel(int no, int nv, const double* __restrict__ eo,
                                                           const double* __restrict__ ev, const double* __restrict__ r1,
                                                           const double* __restrict__ r2, double* t1, double* t2,
                                                           double* scratch) {
  __shared__ double red[8];
  const int vv = nv * nv;
  double part = 0.0;
  for (int ij = blockIdx.x; ij < no * no; ij += gridDim.x) {
    const int i = ij / no, j = ij - i * no;
    const double eij = eo[i] + eo[j];
    const i64 base = (i64)ij * vv;
    for (int ab = threadIdx.x; ab < vv; ab += blockDim.x) {
      const int a = ab / nv, b = ab - a * nv;
      const double d = r2[base + ab] / (eij - ev[a] - ev[b]);
      t2[base + ab] += d;
      part += d * d;
    }
  }
  if (blockIdx.x == 0 && r1 != nullptr) {
    for (int e = threadIdx.x; e < no * nv; e += blockDim.x) {
      const int i = e / nv, a = e - i * nv;
      const double d = r1[e] / (eo[i] - ev[a]);
      t1[e] += d;
      part += d * d;
    }
  }
  part = block_sum<256>(part, red);
  if (threadIdx.x == 0) scratch[blockIdx.x] = part;
}

// ---- E = 2 f.t1 + (t2 + t1 t1).L --------------------------------------------------------------------
__global__ void __launch_bounds__(256) energy_kernel(int no, int nv, const double* __restrict__ fov, i64 ldf,
                                                     const double* __restrict__ t1, const double* __restrict__ t2,
                                                     const double* __restrict__ L, double* scratch) {
  __shared__ double red[8];
  const int vv = nv * nv;
  double part = 0.0;
  for (int ij = blockIdx.x; ij < no * no; ij += gridDim.x) {
    const int i = ij / no, j = ij - i * no;
    const double* ti = t1 + (i64)i * nv;
    const double* tj = t1 + (i64)j * nv;
    const i64 base = (i64)ij * vv;
    for (int ab = threadIdx.x; ab < vv; ab += blockDim.x) {
      const int a = ab / nv, b = ab - a * nv;
      part += (t2[base + ab] + ti[a] * tj[b]) * L[base + ab];
    }
  }
  if (blockIdx.x == 0) {
    for (int e = threadIdx.x; e < no * nv; e += blockDim.x) {
      const int i = e / nv, a = e - i * nv;
      part += 2.0 * fov[(i64)i * ldf + a] * t1[e];
    }
  }
  part = block_sum<256>(part, red);
  if (threadIdx.x == 0) scratch[blockIdx.x] = part;
}

// ---- the same two passes for a RANGE OF ROWS i in [i0, i1) (multi-GPU: each rank updates only its rows of t2 after the
// all-reduce of the half residual, then the rows are all-gathered; parallel.py) --------------------------------------
// r[i,j,a,b] = half[i,j,a,b] + half[j,i,b,a];  t2[i,j,a,b] += r/D;  partial sum of (r/D)^2.  No partner writes: the
// rank that owns row j does (j,i) itself.  Work item = (i, j, tile A, tile B); block 32 x 8.
__global__ void __launch_bounds__(256) update_rows_kernel(int no, int nv, int i0, int i1, const double* __restrict__ eo,
                                                          const double* __restrict__ ev, const double* __restrict__ half,
                                                          double* t2, double* scratch) {
  __shared__ double Y[32][33];
  __shared__ double red[8];
  const int nt = (nv + 31) / 32;
  const i64 nwork = (i64)(i1 - i0) * no * nt * nt;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const i64 vv = (i64)nv * nv;
  double part = 0.0;
  for (i64 w = blockIdx.x; w < nwork; w += gridDim.x) {
    i64 r = w;
    const int tb = (int)(r % nt); r /= nt;
    const int ta = (int)(r % nt); r /= nt;
    const int j = (int)(r % no);
    const int i = i0 + (int)(r / no);
    const int a0 = ta * 32, b0 = tb * 32;
    const double* Hij = half + ((i64)i * no + j) * vv;
    const double* Hji = half + ((i64)j * no + i) * vv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const int y = ty + k;       // Y[y][x] = H_ji[b0+y][a0+x]
      Y[y][tx] = (b0 + y < nv && a0 + tx < nv) ? Hji[(i64)(b0 + y) * nv + a0 + tx] : 0.0;
    }
    __syncthreads();
    const double eij = eo[i] + eo[j];
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      const int y = ty + k;
      const int a = a0 + y, b = b0 + tx;
      if (a < nv && b < nv) {
        const i64 off = ((i64)i * no + j) * vv + (i64)a * nv + b;
        const double d = (Hij[(i64)a * nv + b] + Y[tx][y]) / (eij - ev[a] - ev[b]);
        t2[off] += d;
        part += d * d;
      }
    }
  }
  __syncthreads();
  {
    const int tid = ty * 32 + tx;
    double v = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
      double sm = tid < 8 ? red[tid] : 0.0;
      sm = warp_sum(sm);
      if (tid == 0) scratch[blockIdx.x] = sm;
    }
  }
}

// t1 += r1/D and sum (r1/D)^2 (tiny; one block)
__global__ void __launch_bounds__(256) update_t1_kernel(int no, int nv, const double* __restrict__ eo,
                                                        const double* __restrict__ ev, const double* __restrict__ r1,
                                                        double* t1, double* out) {
  __shared__ double red[8];
  double part = 0.0;
  for (int e = threadIdx.x; e < no * nv; e += 256) {
    const int i = e / nv, a = e - i * nv;
    const double d = r1[e] / (eo[i] - ev[a]);
    t1[e] += d;
    part += d * d;
  }
  part = block_sum<256>(part, red);
  if (threadIdx.x == 0) out[0] = part;
}

__global__ void __launch_bounds__(256) energy_rows_kernel(int no, int nv, int i0, int i1, int singles,
                                                          const double* __restrict__ fov, i64 ldf,
                                                          const double* __restrict__ t1, const double* __restrict__ t2,
                                                          const double* __restrict__ L, double* scratch) {
  __shared__ double red[8];
  const int vv = nv * nv;
  double part = 0.0;
  for (int ij = i0 * no + blockIdx.x; ij < i1 * no; ij += gridDim.x) {
    const int i = ij / no, j = ij - i * no;
    const double* ti = t1 + (i64)i * nv;
    const double* tj = t1 + (i64)j * nv;
    const i64 base = (i64)ij * vv;
    for (int ab = threadIdx.x; ab < vv; ab += blockDim.x) {
      const int a = ab / nv, b = ab - a * nv;
      part += (t2[base + ab] + ti[a] * tj[b]) * L[base + ab];
    }
  }
  if (blockIdx.x == 0 && singles) {
    for (int e = threadIdx.x; e < no * nv; e += blockDim.x) {
      const int i = e / nv, a = e - i * nv;
      part += 2.0 * fov[(i64)i * ldf + a] * t1[e];
    }
  }
  part = block_sum<256>(part, red);
  if (threadIdx.x == 0) scratch[blockIdx.x] = part;
}

// ---- DIIS helpers ------------------------------------------------------------------------------------
struct PtrPack {
  const double* p[16];
  double c[16];
};

template <int M>
__global__ void __launch_bounds__(256) multi_dot_kernel(i64 n, const double* __restrict__ x, PtrPack ys,
                                                        double* scratch, int stride) {
  __shared__ double red[8];
  double acc[M];
#pragma unroll
  for (int q = 0; q < M; ++q) acc[q] = 0.0;
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x) {
    const double xv = x[e];
#pragma unroll
    for (int q = 0; q < M; ++q) acc[q] += xv * ys.p[q][e];
  }
#pragma unroll
  for (int q = 0; q < M; ++q) {
    const double s = block_sum<256>(acc[q], red);
    if (threadIdx.x == 0) scratch[(i64)q * stride + blockIdx.x] = s;
  }
}

template <int M>
__global__ void __launch_bounds__(256) multi_axpy_kernel(i64 n, PtrPack xs, double* __restrict__ out) {
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < M; ++q) v += xs.c[q] * xs.p[q][e];
    out[e] = v;
  }
}

}  // namespace b200cc

using namespace b200cc;

extern "C" int b200cc_version(void) { return 200; }
extern "C" const char* b200cc_last_error(void) { return g_err; }
extern "C" b200cc_i64 b200cc_launch_count(void) { return g_launches.load(); }

extern "C" int b200cc_device_info(int* sms, int* cc_major, int* cc_minor, b200cc_i64* free_bytes,
                                  b200cc_i64* total_bytes) {
  int dev = 0;
  B200CC_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  B200CC_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  size_t f = 0, t = 0;
  B200CC_CUDA_OK(cudaMemGetInfo(&f, &t));
  if (sms) *sms = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (free_bytes) *free_bytes = (b200cc_i64)f;
  if (total_bytes) *total_bytes = (b200cc_i64)t;
  return 0;
}

static int grid_rows(int rows) {
  const int cap = sm_count() * 8;
  return rows < cap ? (rows > 0 ? rows : 1) : cap;
}

extern "C" int b200cc_build_tau(int no, int nv, double f1, double f2, const double* t1, const double* t2,
                                double* tau, void* stream) {
  if (no <= 0 || nv <= 0) return 0;
  tau_kernel<<<grid_rows(no * no), 256, 0, static_cast<cudaStream_t>(stream)>>>(no, nv, f1, f2, t1, t2, tau);
  return check_launch("tau_kernel");
}

extern "C" int b200cc_div_d2(int no, int nv, const double* eo, const double* ev, const double* in, double* out,
                             void* stream) {
  if (no <= 0 || nv <= 0) return 0;
  div_d2_kernel<<<grid_rows(no * no), 256, 0, static_cast<cudaStream_t>(stream)>>>(no, nv, eo, ev, in, out);
  return check_launch("div_d2_kernel");
}

extern "C" int b200cc_div_d1(int no, int nv, const double* eo, const double* ev, const double* in, double* out,
                             void* stream) {
  if (no <= 0 || nv <= 0) return 0;
  const int n = no * nv;
  div_d1_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(no, nv, eo, ev, in, out);
  return check_launch("div_d1_kernel");
}

extern "C" int b200cc_update_amps(int no, int nv, const double* eo, const double* ev, const double* r1,
                                  double* r2_half, int symmetrize, int write_r2, double* t1, double* t2,
                                  double* sumsq, double* scratch, void* stream) {
  if (no <= 0 || nv <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int nblk;
  if (symmetrize) {
    const int nt = (nv + 31) / 32;
    const i64 nwork = (i64)no * (no + 1) / 2 * nt * nt;
    nblk = (int)(nwork < 2048 ? nwork : 2048);
    if (write_r2)
      sym_update_kernel<true><<<nblk, dim3(32, 8), 0, st>>>(no, nv, eo, ev, r1, r2_half, t1, t2, scratch, 1);
    else
      sym_update_kernel<false><<<nblk, dim3(32, 8), 0, st>>>(no, nv, eo, ev, r1, r2_half, t1, t2, scratch, 1);
    if (check_launch("sym_update_kernel")) return 1;
  } else {
    nblk = no * no < 2048 ? no * no : 2048;
    update_plain_kernel<<<nblk, 256, 0, st>>>(no, nv, eo, ev, r1, r2_half, t1, t2, scratch);
    if (check_launch("update_plain_kernel")) return 1;
  }
  return launch_final_reduce(scratch, nblk, 0, 1, sumsq, 0, 1.0, st);
}

extern "C" int b200cc_symmetrize_r2(int no, int nv, double* r2, void* stream) {
  if (no <= 0 || nv <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nt = (nv + 31) / 32;
  const i64 nwork = (i64)no * (no + 1) / 2 * nt * nt;
  const int nblk = (int)(nwork < 4096 ? nwork : 4096);
  // scratch-free variant: reuse the fused kernel without update; partials go to a static device buffer
  static double* dummy = nullptr;
  if (!dummy) B200CC_CUDA_OK(cudaMalloc(&dummy, 4096 * sizeof(double)));
  sym_update_kernel<true><<<nblk, dim3(32, 8), 0, st>>>(no, nv, nullptr, nullptr, nullptr, r2, nullptr, nullptr,
                                                       dummy, 0);
  return check_launch("sym_update_kernel(symmetrize)");
}

extern "C" int b200cc_cc_energy(int no, int nv, const double* fov, b200cc_i64 ldf, const double* t1,
                                const double* t2, const double* Loovv, double* e_out, double* scratch,
                                void* stream) {
  if (no <= 0 || nv <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = no * no < 2048 ? no * no : 2048;
  energy_kernel<<<nblk, 256, 0, st>>>(no, nv, fov, ldf, t1, t2, Loovv, scratch);
  if (check_launch("energy_kernel")) return 1;
  return launch_final_reduce(scratch, nblk, 0, 1, e_out, 0, 1.0, st);
}

extern "C" int b200cc_update_amps_rows(int no, int nv, int i0, int i1, const double* eo, const double* ev,
                                       const double* r1, const double* r2_half, double* t1, double* t2,
                                       double* sumsq2, double* scratch, void* stream) {
  if (no <= 0 || nv <= 0) return 0;
  if (i0 < 0 || i1 > no || i1 < i0) { set_error("b200cc_update_amps_rows: bad row range"); return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // sumsq2[0] = doubles of rows [i0,i1), sumsq2[1] = singles (every caller updates the whole, replicated t1)
  B200CC_CUDA_OK(cudaMemsetAsync(sumsq2, 0, 2 * sizeof(double), st));
  if (i1 > i0) {
    const int nt = (nv + 31) / 32;
    const i64 nwork = (i64)(i1 - i0) * no * nt * nt;
    const int nblk = (int)(nwork < 2048 ? nwork : 2048);
    update_rows_kernel<<<nblk, dim3(32, 8), 0, st>>>(no, nv, i0, i1, eo, ev, r2_half, t2, scratch);
    if (check_launch("update_rows_kernel")) return 1;
    if (launch_final_reduce(scratch, nblk, 0, 1, sumsq2, 0, 1.0, st)) return 1;
  }
  if (r1 != nullptr) {
    update_t1_kernel<<<1, 256, 0, st>>>(no, nv, eo, ev, r1, t1, sumsq2 + 1);
    if (check_launch("update_t1_kernel")) return 1;
  }
  return 0;
}

extern "C" int b200cc_cc_energy_rows(int no, int nv, int i0, int i1, int with_singles, const double* fov,
                                     b200cc_i64 ldf, const double* t1, const double* t2, const double* Loovv,
                                     double* e_out, double* scratch, void* stream) {
  if (no <= 0 || nv <= 0) return 0;
  if (i0 < 0 || i1 > no || i1 < i0) { set_error("b200cc_cc_energy_rows: bad row range"); return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = (i1 - i0) * no;
  const int nblk = rows < 1 ? 1 : (rows < 2048 ? rows : 2048);
  energy_rows_kernel<<<nblk, 256, 0, st>>>(no, nv, i0, i1, with_singles, fov, ldf, t1, t2, Loovv, scratch);
  if (check_launch("energy_rows_kernel")) return 1;
  return launch_final_reduce(scratch, nblk, 0, 1, e_out, 0, 1.0, st);
}

extern "C" int b200cc_multi_dot(b200cc_i64 n, const double* x, int m, const double* const* ys, double* out,
                                double* scratch, void* stream) {
  if (m < 1 || m > 16) { set_error("b200cc_multi_dot: m=%d not in [1,16]", m); return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PtrPack pk;
  for (int q = 0; q < 16; ++q) { pk.p[q] = ys[q < m ? q : 0]; pk.c[q] = 0.0; }
  i64 want = (n + 255) / 256;
  const int nblk = (int)(want < 1 ? 1 : (want < 1024 ? want : 1024));
  const int stride = 1024;
  switch (m) {
#define B200CC_CASE(M) \
  case M: multi_dot_kernel<M><<<nblk, 256, 0, st>>>(n, x, pk, scratch, stride); break;
    B200CC_CASE(1) B200CC_CASE(2) B200CC_CASE(3) B200CC_CASE(4) B200CC_CASE(5) B200CC_CASE(6) B200CC_CASE(7)
    B200CC_CASE(8) B200CC_CASE(9) B200CC_CASE(10) B200CC_CASE(11) B200CC_CASE(12) B200CC_CASE(13)
    B200CC_CASE(14) B200CC_CASE(15) B200CC_CASE(16)
#undef B200CC_CASE
  }
  if (check_launch("multi_dot_kernel")) return 1;
  return launch_final_reduce(scratch, nblk, stride, m, out, 0, 1.0, st);
}

extern "C" int b200cc_multi_axpy(b200cc_i64 n, int m, const double* c, const double* const* xs, double* out,
                                 void* stream) {
  if (m < 1 || m > 16) { set_error("b200cc_multi_axpy: m=%d not in [1,16]", m); return 1; }
  if (n <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PtrPack pk;
  for (int q = 0; q < 16; ++q) { pk.p[q] = xs[q < m ? q : 0]; pk.c[q] = q < m ? c[q] : 0.0; }
  i64 want = (n + 255) / 256;
  const int cap = sm_count() * 16;
  const int nblk = (int)(want < cap ? want : cap);
  switch (m) {
#define B200CC_CASE(M) \
  case M: multi_axpy_kernel<M><<<nblk, 256, 0, st>>>(n, pk, out); break;
    B200CC_CASE(1) B200CC_CASE(2) B200CC_CASE(3) B200CC_CASE(4) B200CC_CASE(5) B200CC_CASE(6) B200CC_CASE(7)
    B200CC_CASE(8) B200CC_CASE(9) B200CC_CASE(10) B200CC_CASE(11) B200CC_CASE(12) B200CC_CASE(13)
    B200CC_CASE(14) B200CC_CASE(15) B200CC_CASE(16)
#undef B200CC_CASE
  }
  return check_launch("multi_axpy_kernel");
}
