// permute.cu -- strided tensor permutation / axpby, HBM-bound.
//   out[sum_d i_d*so[d]] = alpha * in[sum_d i_d*si[d]] + beta * out[...]
// Two kernels:
//   rowcopy  : the fastest-varying loop dimension is (near-)contiguous on both sides; one thread per
//              element, consecutive threads along that dimension (coalesced loads and stores).
//   transpose: input and output are contiguous along DIFFERENT dimensions; a 32x32 tile of those two
//              dimensions goes through shared memory (33-column padding) so both sides are coalesced.
#include "common.cuh"

namespace b200cc {

constexpr int MAXR = 6;

struct PermParams {
  int rank;
  i64 shape[MAXR], si[MAXR], so[MAXR];
  double alpha, beta;
  i64 total;
  // transpose kernel: dx = dimension contiguous in the input, dy = dimension contiguous in the output
  int dx, dy;
  i64 tiles_x, tiles_y, outer;
  int width;   // rowcopy: threads that share one row (a divisor of 256; chosen by the host to waste the fewest lanes)
};

// Row length n handled by groups of `width` threads in ceil(n / width) passes: pick the width in {256,128,64,32} (or n
// itself below 256... rounded down to a divisor of 256 is not needed: short rows use width = n) that idles the fewest
// lanes -- v = 300 on 256 lanes leaves 41 % of the lane-slots empty (measured 2.6-2.9 TB/s), on 64 lanes 6 %.
static int rowcopy_width(i64 n) {
  if (n < 256) return (int)n;
  int best = 256;
  double waste = (double)((n + 255) / 256 * 256) / (double)n;
  for (int w = 128; w >= 32; w >>= 1) {
    const double x = (double)((n + w - 1) / w * w) / (double)n;
    if (x < waste - 0.02) { waste = x; best = w; }
  }
  return best;
}

// One CTA per "row" (all dims but the last, grid-stride), threads along the last dimension: the 64-bit
// div/mod index decode happens once per row per thread instead of once per element.
__global__ void __launch_bounds__(256) permute_rowcopy_kernel(const PermParams p, const double* __restrict__ in,
                                                              double* __restrict__ out) {
  const i64 n_last = p.shape[p.rank - 1];
  const i64 si_l = p.si[p.rank - 1], so_l = p.so[p.rank - 1];
  if (p.rank == 1) {   // plain strided vector: grid-stride over the elements
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n_last; e += (i64)gridDim.x * blockDim.x) {
      const double v = p.alpha * in[e * si_l];
      out[e * so_l] = (p.beta != 0.0) ? v + p.beta * out[e * so_l] : v;
    }
    return;
  }
  const i64 rows = p.total / n_last;
  // each CTA owns a contiguous range of rows; short rows are processed rpc at a time so the CTA stays full
  const int width = p.width;
  const int rpc = 256 / width;
  const int sub = threadIdx.x / width;
  const i64 lane0 = threadIdx.x - sub * width;
  if (sub >= rpc) return;
  const i64 per = (rows + gridDim.x - 1) / gridDim.x;
  const i64 r_end = min(rows, ((i64)blockIdx.x + 1) * per);
  i64 row = (i64)blockIdx.x * per + sub;
  if (row >= r_end) return;
  // decode the first row once, then advance the multi-index like an odometer (no div/mod per row)
  i64 idx[MAXR - 1];
  {
    i64 r = row;
#pragma unroll
    for (int d = MAXR - 2; d >= 0; --d) {
      idx[d] = 0;
      if (d < p.rank - 1) {
        const i64 s = p.shape[d];
        const i64 qd = r / s;
        idx[d] = r - qd * s;
        r = qd;
      }
    }
  }
  // UNR rows per thread per iteration: all loads are issued before the stores (bytes in flight, not index
  // arithmetic, bound this kernel: one 8-byte load per thread in flight gave 2.0 TB/s)
  constexpr int UNR = 4;
  for (; row < r_end; row += (i64)UNR * rpc) {
    i64 oi[UNR], oo[UNR];
    bool ok[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      ok[u] = row + (i64)u * rpc < r_end;
      oi[u] = 0;
      oo[u] = 0;
#pragma unroll
      for (int d = 0; d < MAXR - 1; ++d) {
        if (d < p.rank - 1) {
          oi[u] += idx[d] * p.si[d];
          oo[u] += idx[d] * p.so[d];
        }
      }
      i64 carry = rpc;
#pragma unroll
      for (int d = MAXR - 2; d >= 0; --d) {
        if (d < p.rank - 1 && carry != 0) {
          idx[d] += carry;
          carry = 0;
          const i64 s = p.shape[d];
          if (idx[d] >= s) {
            carry = idx[d] / s;
            idx[d] -= carry * s;
          }
        }
      }
    }
    for (i64 e = lane0; e < n_last; e += width) {
      double v[UNR], w[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) v[u] = ok[u] ? in[oi[u] + e * si_l] : 0.0;
      if (p.beta != 0.0) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) w[u] = ok[u] ? out[oo[u] + e * so_l] : 0.0;
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          if (ok[u]) out[oo[u] + e * so_l] = p.alpha * v[u] + p.beta * w[u];
      } else {
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          if (ok[u]) out[oo[u] + e * so_l] = p.alpha * v[u];
      }
    }
  }
}

// grid.x enumerates (outer index, tile_y, tile_x); block = 32 x 8
__global__ void __launch_bounds__(256) permute_transpose_kernel(const PermParams p, const double* __restrict__ in,
                                                                double* __restrict__ out) {
  __shared__ double tile[32][33];
  i64 bid = blockIdx.x;
  const i64 tx = bid % p.tiles_x; bid /= p.tiles_x;
  const i64 ty = bid % p.tiles_y; bid /= p.tiles_y;
  // remaining dims (all except dx, dy) from bid
  i64 oi = 0, oo = 0, r = bid;
#pragma unroll
  for (int d = MAXR - 1; d >= 0; --d) {
    if (d < p.rank && d != p.dx && d != p.dy) {
      const i64 s = p.shape[d];
      const i64 qd = r / s;
      const i64 id = r - qd * s;
      r = qd;
      oi += id * p.si[d];
      oo += id * p.so[d];
    }
  }
  const i64 x0 = tx * 32, y0 = ty * 32;
  const i64 nx = p.shape[p.dx], ny = p.shape[p.dy];
  const i64 six = p.si[p.dx], siy = p.si[p.dy], sox = p.so[p.dx], soy = p.so[p.dy];
  // load: threadIdx.x runs along dx (input-contiguous)
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const i64 x = x0 + threadIdx.x, y = y0 + threadIdx.y + k;
    if (x < nx && y < ny) tile[threadIdx.y + k][threadIdx.x] = in[oi + x * six + y * siy];
  }
  __syncthreads();
  // store: threadIdx.x runs along dy (output-contiguous)
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const i64 y = y0 + threadIdx.x, x = x0 + threadIdx.y + k;
    if (x < nx && y < ny) {
      const i64 o = oo + x * sox + y * soy;
      const double v = p.alpha * tile[threadIdx.x][threadIdx.y + k];
      out[o] = (p.beta != 0.0) ? v + p.beta * out[o] : v;
    }
  }
}

__global__ void __launch_bounds__(256) axpbyz_kernel(i64 n, double a, const double* x, double b, const double* y,
                                                     double* z) {
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x) {
    double v = 0.0;
    if (a != 0.0) v += a * x[e];
    if (b != 0.0) v += b * y[e];
    z[e] = v;
  }
}

static i64 iabs(i64 x) { return x < 0 ? -x : x; }

}  // namespace b200cc

using namespace b200cc;

extern "C" int b200cc_permute(int rank, const b200cc_i64* shape, const b200cc_i64* si, const b200cc_i64* so,
                              double alpha, const double* in, double beta, double* out, void* stream) {
  if (rank < 0 || rank > MAXR) { set_error("b200cc_permute: rank %d not in [0,%d]", rank, MAXR); return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PermParams p;
  // drop extent-1 dims, detect empty
  p.rank = 0;
  p.total = 1;
  for (int d = 0; d < rank; ++d) {
    if (shape[d] < 0) { set_error("b200cc_permute: negative extent"); return 1; }
    if (shape[d] == 0) return 0;
    if (shape[d] == 1) continue;
    p.shape[p.rank] = shape[d]; p.si[p.rank] = si[d]; p.so[p.rank] = so[d];
    p.total *= shape[d];
    ++p.rank;
  }
  if (p.rank == 0) { p.rank = 1; p.shape[0] = 1; p.si[0] = 0; p.so[0] = 0; }
  // sort loop dims by DEcreasing |output stride| so that the last dim is the output-fastest one
  for (int a = 0; a < p.rank; ++a)
    for (int b = a + 1; b < p.rank; ++b)
      if (iabs(p.so[b]) > iabs(p.so[a])) {
        i64 t;
        t = p.shape[a]; p.shape[a] = p.shape[b]; p.shape[b] = t;
        t = p.si[a]; p.si[a] = p.si[b]; p.si[b] = t;
        t = p.so[a]; p.so[a] = p.so[b]; p.so[b] = t;
      }
  for (int d = p.rank; d < MAXR; ++d) { p.shape[d] = 1; p.si[d] = 0; p.so[d] = 0; }
  p.alpha = alpha; p.beta = beta;
  const int last = p.rank - 1;
  // input-fastest dim
  int dx = 0;
  for (int d = 1; d < p.rank; ++d)
    if (p.si[d] != 0 && (p.si[dx] == 0 || iabs(p.si[d]) < iabs(p.si[dx]))) dx = d;
  const bool need_transpose = p.rank >= 2 && dx != last && iabs(p.si[last]) != 1 && p.si[dx] != 0 &&
                              p.shape[dx] >= 8 && p.shape[last] >= 8;
  const int cap = sm_count() * 16;
  if (!need_transpose) {
    const i64 n_last = p.shape[last];
    p.width = rowcopy_width(n_last);
    const i64 rpc = 256 / p.width;
    i64 blocks = p.rank == 1 ? (n_last + 255) / 256 : (p.total / n_last + 16 * rpc - 1) / (16 * rpc);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    permute_rowcopy_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, in, out);
    return check_launch("permute_rowcopy_kernel");
  }
  p.dx = dx; p.dy = last;
  p.tiles_x = (p.shape[dx] + 31) / 32;
  p.tiles_y = (p.shape[last] + 31) / 32;
  p.outer = p.total / (p.shape[dx] * p.shape[last]);
  const i64 nblocks = p.tiles_x * p.tiles_y * p.outer;
  if (nblocks > 2147483647LL) { set_error("b200cc_permute: grid too large"); return 1; }
  permute_transpose_kernel<<<(unsigned)nblocks, dim3(32, 8), 0, st>>>(p, in, out);
  return check_launch("permute_transpose_kernel");
}

extern "C" int b200cc_axpbyz(b200cc_i64 n, double a, const double* x, double b, const double* y, double* z,
                             void* stream) {
  if (n <= 0) return 0;
  i64 blocks = (n + 255) / 256;
  const int cap = sm_count() * 16;
  if (blocks > cap) blocks = cap;
  axpbyz_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(n, a, x, b, y, z);
  return check_launch("axpbyz_kernel");
}
