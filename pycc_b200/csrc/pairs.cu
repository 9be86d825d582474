// pairs.cu -- the symmetric / antisymmetric ("packed pair") form of the particle-particle ladder
//   r2[i,j,a,b] += alpha * sum_ef tau[i,j,e,f] <ab|ef>                                   (pycc/ccwfn.py:931)
//
// <ab|ef> = <ba|fe> (real orbitals), so with the pair indices P = (a >= b), Q = (e >= f), pair(x,y) = x(x+1)/2 + y,
//   V+[P,Q] = <ab|ef> + <ab|fe>  (e > f),  <ab|ee>  (e = f)          symmetric under a <-> b
//   V-[P,Q] = <ab|ef> - <ab|fe>                                       antisymmetric under a <-> b (zero for a = b or e = f)
//   T+[m,Q] = 1/2 (tau[m,e,f] + tau[m,f,e])  (tau[m,e,e] for e = f),   T-[m,Q] = 1/2 (tau[m,e,f] - tau[m,f,e])
// the ladder is  S = T+ V+^T,  A = T- V-^T  (two GEMMs with N = K = v(v+1)/2 instead of one with N = K = v^2) and
//   L[m,a,b] = S[m,P] + A[m,P],   L[m,b,a] = S[m,P] - A[m,P].
// This is the split the reference itself writes for the CC3 intermediate W_abei (pycc/ccwfn.py:1054-1120).  If tau has
// the pair symmetry tau[i,j,e,f] = tau[j,i,f,e] (true for the amplitudes solve_cc iterates), T+ is symmetric and T-
// antisymmetric under i <-> j as well, so only rows m = (i >= j) are needed ("tri" mode): executed flops o^2 v^4 / 2
// instead of 2 o^2 v^4, and <ab|ef> is held as V+ and V- (v^4 / 2 doubles instead of v^4).
//
// Kernels here are the HBM-bound passes around that GEMM: packing <ab|ef> rows into V+/V- (once), packing tau into T+/T-
// and scattering S/A back into r2 (every iteration), and rebuilding FP64 rows of <ab|ef> from V+/V- for the few other
// consumers of the block.
#include "common.cuh"

namespace b200cc {

__host__ __device__ __forceinline__ i64 pair_index(i64 x, i64 y) { return x * (x + 1) / 2 + y; }

// largest x with x(x+1)/2 <= p
__device__ __forceinline__ int pair_row(i64 p) {
  int x = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
  while (pair_index(x + 1, 0) <= p) ++x;
  while (pair_index(x, 0) > p) --x;
  return x;
}

// ---- <ab|ef> rows -> V+, V- ----------------------------------------------------------------------------------------
// one CTA per pair (a,b), a in [a0,a1), b <= a; slab(a,b)[e,f] = src[(a-a0)*sa + b*sb + e*se + f*sf]
__global__ void __launch_bounds__(256) pack_pairs_kernel(const double* __restrict__ src, i64 sa, i64 sb, i64 se, i64 sf,
                                                         int nv, int a0, i64 npairs, double* __restrict__ vp,
                                                         double* __restrict__ vm, i64 ldq) {
  const i64 nq = pair_index(nv, 0);
  const i64 pbase = pair_index(a0, 0);
  for (i64 pl = blockIdx.x; pl < npairs; pl += gridDim.x) {
    const int a = pair_row(pl + pbase);
    const int b = (int)(pl + pbase - pair_index(a, 0));
    const double* S = src + (i64)(a - a0) * sa + (i64)b * sb;
    double* op = vp + pl * ldq;
    double* om = vm + pl * ldq;
    for (int e = 0; e < nv; ++e) {
      const i64 q0 = pair_index(e, 0);
      for (int f = threadIdx.x; f <= e; f += blockDim.x) {
        const double x = S[(i64)e * se + (i64)f * sf];
        const double y = S[(i64)f * se + (i64)e * sf];
        const bool dg = (f == e);
        op[q0 + f] = dg ? x : x + y;
        om[q0 + f] = (dg || a == b) ? 0.0 : x - y;
      }
    }
    for (i64 q = nq + threadIdx.x; q < ldq; q += blockDim.x) { op[q] = 0.0; om[q] = 0.0; }
  }
}

// ---- V+, V- rows -> FP64 <ab|ef> slabs: dst[p][e][f], p = 0 .. npairs-1 -----------------------------------------------
__global__ void __launch_bounds__(256) unpack_pairs_kernel(const double* __restrict__ vp, const double* __restrict__ vm,
                                                           i64 ldq, int nv, i64 npairs, double* __restrict__ dst) {
  const i64 vv = (i64)nv * nv;
  for (i64 pl = blockIdx.x; pl < npairs; pl += gridDim.x) {
    const double* ip = vp + pl * ldq;
    const double* im = vm + pl * ldq;
    double* D = dst + pl * vv;
    for (i64 ef = threadIdx.x; ef < vv; ef += blockDim.x) {
      const int e = (int)(ef / nv), f = (int)(ef - (i64)e * nv);
      double r;
      if (e == f) r = ip[pair_index(e, e)];
      else if (e > f) { const i64 q = pair_index(e, f); r = 0.5 * (ip[q] + im[q]); }
      else { const i64 q = pair_index(f, e); r = 0.5 * (ip[q] - im[q]); }
      D[ef] = r;
    }
  }
}

// ---- tau -> T+, T- --------------------------------------------------------------------------------------------------
// grid.x = 32x32 tile pairs (te >= tf) of the (e,f) plane, grid.y = rows m; block 32 x 8.
// tri = 0: m = i*no + j over all (i,j);  tri = 1: m = pair(i,j) over i >= j, slab tau[i,j].
// fs = 1/2 for amplitudes (T+-), 1 for integral rows (X+- = x_ef +- x_fe, b200cc_pack_rows); nrows: number of (v,v) slabs
// when tri == 0
__global__ void __launch_bounds__(256) pack_tau_kernel(const double* __restrict__ tau, i64 nrows, int no, int nv, int tri,
                                                       double fs, double* __restrict__ tp, double* __restrict__ tm,
                                                       i64 ldq) {
  __shared__ double X[32][33], Y[32][33];
  const int nt = (nv + 31) / 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const i64 vv = (i64)nv * nv;
  const i64 nq = pair_index(nv, 0);
  const int ntp = nt * (nt + 1) / 2;
  for (i64 m = blockIdx.y; m < (tri ? pair_index(no, 0) : nrows); m += gridDim.y) {
    i64 slab = m;
    if (tri) {
      const int i = pair_row(m);
      slab = (i64)i * no + (m - pair_index(i, 0));
    }
    const double* T = tau + slab * vv;
    double* op = tp + m * ldq;
    double* om = tm + m * ldq;
    for (int w = blockIdx.x; w < ntp + 1; w += gridDim.x) {
      if (w == ntp) {                                    // the pad columns of this row
        for (i64 q = nq + ty * 32 + tx; q < ldq; q += 256) { op[q] = 0.0; om[q] = 0.0; }
        continue;
      }
      const int te = pair_row(w), tf = w - (int)pair_index(te, 0);
      const int e0 = te * 32, f0 = tf * 32;
      __syncthreads();
      for (int r = ty; r < 32; r += 8) {
        const int e = e0 + r, f = f0 + tx;
        X[r][tx] = (e < nv && f < nv) ? T[(i64)e * nv + f] : 0.0;          // tau[e, f]
        const int e2 = f0 + r, f2 = e0 + tx;
        Y[r][tx] = (e2 < nv && f2 < nv) ? T[(i64)e2 * nv + f2] : 0.0;      // tau[f-tile row, e-tile col]
      }
      __syncthreads();
      for (int r = ty; r < 32; r += 8) {
        const int e = e0 + r, f = f0 + tx;
        if (e < nv && f <= e) {
          const double x = X[r][tx], y = Y[tx][r];                         // tau[e,f], tau[f,e]
          const i64 q = pair_index(e, f);
          op[q] = (e == f) ? x : fs * (x + y);
          om[q] = (e == f) ? 0.0 : fs * (x - y);
        }
      }
    }
  }
}

// ---- S, A -> r2 -------------------------------------------------------------------------------------------------------
// S, A: [M][lds], column = pair(a,b) - pair(a0,0) for a in [a0,a1), b <= a.
//   r2[i,j,a,b] += alpha (S + A),  r2[i,j,b,a] += alpha (S - A)  (a != b), and in tri mode (rows m = (i >= j), tau pair-symmetric)
//   for i != j also  r2[j,i,a,b] += alpha (S - A),  r2[j,i,b,a] += alpha (S + A).
// grid.x = 32x32 tile pairs (ta >= tb), grid.y = rows m; block 32 x 8; every r2 element is touched by exactly one thread.
__global__ void __launch_bounds__(256) ladder_unpack_kernel(const double* __restrict__ S, const double* __restrict__ A,
                                                            i64 lds, int no, int nv, int tri, int a0, int a1,
                                                            double alpha, double* __restrict__ r2) {
  __shared__ double U[32][33], W[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const i64 vv = (i64)nv * nv;
  const i64 pbase = pair_index(a0, 0);
  const int t_lo = a0 / 32, t_hi = (a1 - 1) / 32;                          // a-tiles that meet [a0, a1)
  const i64 w_lo = pair_index(t_lo, 0), w_hi = pair_index(t_hi + 1, 0);
  for (i64 m = blockIdx.y; m < (tri ? pair_index(no, 0) : (i64)no * no); m += gridDim.y) {
    int i, j;
    if (tri) { i = pair_row(m); j = (int)(m - pair_index(i, 0)); }
    else { i = (int)(m / no); j = (int)(m - (i64)i * no); }
    const double* Sm = S + m * lds;
    const double* Am = A + m * lds;
    double* Rij = r2 + ((i64)i * no + j) * vv;
    double* Rji = r2 + ((i64)j * no + i) * vv;
    const bool both = tri && i != j;
    for (i64 w = w_lo + blockIdx.x; w < w_hi; w += gridDim.x) {
      const int ta = pair_row(w), tb = (int)(w - pair_index(ta, 0));
      const int ab0 = ta * 32, bb0 = tb * 32;
      __syncthreads();
      for (int r = ty; r < 32; r += 8) {
        const int a = ab0 + r, b = bb0 + tx;
        double u = 0.0, v = 0.0;
        if (a >= a0 && a < a1 && b <= a) {
          const i64 p = pair_index(a, b) - pbase;
          const double s = Sm[p], d = Am[p];
          u = alpha * (s + d);
          v = alpha * (s - d);
        }
        U[r][tx] = u;
        W[r][tx] = v;
      }
      __syncthreads();
      for (int r = ty; r < 32; r += 8) {
        {   // direct orientation: element (a, b), b <= a
          const int a = ab0 + r, b = bb0 + tx;
          if (a >= a0 && a < a1 && b <= a) {
            Rij[(i64)a * nv + b] += U[r][tx];
            if (both) Rji[(i64)a * nv + b] += W[r][tx];
          }
        }
        {   // transposed orientation: element (b, a) with b < a; thread (r, tx) handles b = bb0 + r, a = ab0 + tx
          const int b = bb0 + r, a = ab0 + tx;
          if (a >= a0 && a < a1 && b < a) {
            Rij[(i64)b * nv + a] += W[tx][r];
            if (both) Rji[(i64)b * nv + a] += U[tx][r];
          }
        }
      }
    }
  }
}

// ---- ring layouts of t2 in one pass -----------------------------------------------------------------------------
// u[i,a,m,e] = 2 t2[i,m,a,e] - t2[i,m,e,a]   and   tb[i,a,m,e] = t2[i,m,e,a]   from ONE read of t2 (both come from the
// (i,m) slab; the transposed element through a 32x32 shared tile).  Replaces three permute passes per iteration.
// grid.x = 32x32 tiles of the (a,e) plane, grid.y = (i,m) slabs; block 32 x 8.
__global__ void __launch_bounds__(256) ring_layouts_kernel(const double* __restrict__ t2, int no, int nv,
                                                           double* __restrict__ u, double* __restrict__ tb) {
  __shared__ double X[32][33], Y[32][33];
  const int nt = (nv + 31) / 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const i64 vv = (i64)nv * nv;
  for (i64 im = blockIdx.y; im < (i64)no * no; im += gridDim.y) {
    const int i = (int)(im / no), m = (int)(im - (i64)i * no);
    const double* T = t2 + im * vv;
    for (int w = blockIdx.x; w < nt * nt; w += gridDim.x) {
      const int a0 = (w / nt) * 32, e0 = (w % nt) * 32;
      __syncthreads();
      for (int r = ty; r < 32; r += 8) {
        const int a = a0 + r, e = e0 + tx;
        X[r][tx] = (a < nv && e < nv) ? T[(i64)a * nv + e] : 0.0;            // t2[i,m,a,e]
        const int e2 = e0 + r, a2 = a0 + tx;
        Y[r][tx] = (e2 < nv && a2 < nv) ? T[(i64)e2 * nv + a2] : 0.0;        // t2[i,m,e,a] with e = e0 + r
      }
      __syncthreads();
      for (int r = ty; r < 32; r += 8) {
        const int a = a0 + r, e = e0 + tx;
        if (a < nv && e < nv) {
          const i64 o = (((i64)i * nv + a) * no + m) * nv + e;
          const double x = X[r][tx], y = Y[tx][r];
          u[o] = 2.0 * x - y;
          tb[o] = y;
        }
      }
    }
  }
}

}  // namespace b200cc

using namespace b200cc;

extern "C" int b200cc_ring_layouts(const double* t2, int no, int nv, double* u, double* tb, void* stream) {
  if (no <= 0 || nv <= 0) return 0;
  const int nt = (nv + 31) / 32;
  const i64 slabs = (i64)no * no;
  dim3 grid((unsigned)(nt * nt), (unsigned)(slabs < 65535 ? slabs : 65535)), block(32, 8);
  ring_layouts_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(t2, no, nv, u, tb);
  return check_launch("ring_layouts_kernel");
}

extern "C" b200cc_i64 b200cc_pair_count(int n) { return pair_index(n, 0); }

extern "C" int b200cc_pack_pairs(const double* src, b200cc_i64 sa, b200cc_i64 sb, b200cc_i64 se, b200cc_i64 sf, int nv,
                                 int a0, int a1, double* vp, double* vm, b200cc_i64 ldq, void* stream) {
  if (nv <= 0 || a0 < 0 || a1 > nv || a1 < a0) { set_error("b200cc_pack_pairs: bad range"); return 1; }
  if (ldq < pair_index(nv, 0)) { set_error("b200cc_pack_pairs: ldq < v(v+1)/2"); return 1; }
  const i64 npairs = pair_index(a1, 0) - pair_index(a0, 0);
  if (npairs == 0) return 0;
  const int grid = (int)(npairs < (i64)sm_count() * 16 ? npairs : (i64)sm_count() * 16);
  pack_pairs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, sa, sb, se, sf, nv, a0, npairs, vp, vm, ldq);
  return check_launch("pack_pairs_kernel");
}

extern "C" int b200cc_unpack_pairs(const double* vp, const double* vm, b200cc_i64 ldq, int nv, b200cc_i64 npairs,
                                   double* dst, void* stream) {
  if (nv <= 0 || npairs < 0) { set_error("b200cc_unpack_pairs: bad size"); return 1; }
  if (npairs == 0) return 0;
  const int grid = (int)(npairs < (i64)sm_count() * 16 ? npairs : (i64)sm_count() * 16);
  unpack_pairs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(vp, vm, ldq, nv, npairs, dst);
  return check_launch("unpack_pairs_kernel");
}

extern "C" int b200cc_pack_tau(const double* tau, int no, int nv, int tri, double* tp, double* tm, b200cc_i64 ldq,
                               void* stream) {
  if (no <= 0 || nv <= 0) { set_error("b200cc_pack_tau: bad size"); return 1; }
  if (ldq < pair_index(nv, 0)) { set_error("b200cc_pack_tau: ldq < v(v+1)/2"); return 1; }
  const int nt = (nv + 31) / 32;
  const i64 M = tri ? pair_index(no, 0) : (i64)no * no;
  dim3 grid((unsigned)(nt * (nt + 1) / 2 + 1), (unsigned)(M < 65535 ? M : 65535)), block(32, 8);
  pack_tau_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(tau, (i64)no * no, no, nv, tri ? 1 : 0, 0.5, tp,
                                                                         tm, ldq);
  return check_launch("pack_tau_kernel");
}

extern "C" int b200cc_pack_rows(const double* src, b200cc_i64 nrows, int nv, double* xp, double* xm, b200cc_i64 ldq,
                                void* stream) {
  if (nrows < 0 || nv <= 0) { set_error("b200cc_pack_rows: bad size"); return 1; }
  if (ldq < pair_index(nv, 0)) { set_error("b200cc_pack_rows: ldq < v(v+1)/2"); return 1; }
  if (nrows == 0) return 0;
  const int nt = (nv + 31) / 32;
  dim3 grid((unsigned)(nt * (nt + 1) / 2 + 1), (unsigned)(nrows < 65535 ? nrows : 65535)), block(32, 8);
  pack_tau_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(src, nrows, 1, nv, 0, 1.0, xp, xm, ldq);
  return check_launch("pack_tau_kernel(rows)");
}

// out[(i*no + j)*ldo + c] = S[p,c] + A[p,c],  out[(j*no + i)*ldo + c] = S[p,c] - A[p,c],  p = pair(i,j), i >= j, c < ncols
namespace b200cc {
__global__ void __launch_bounds__(256) pair_rows_unpack_kernel(const double* __restrict__ S, const double* __restrict__ A,
                                                               i64 lds, int no, i64 ncols, double* __restrict__ out,
                                                               i64 ldo) {
  const i64 np = pair_index(no, 0);
  for (i64 p = blockIdx.y; p < np; p += gridDim.y) {
    const int i = pair_row(p), j = (int)(p - pair_index(i, 0));
    const double* Sp = S + p * lds;
    const double* Ap = A + p * lds;
    double* oij = out + ((i64)i * no + j) * ldo;
    double* oji = out + ((i64)j * no + i) * ldo;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < ncols; c += (i64)gridDim.x * blockDim.x) {
      const double s = Sp[c], d = Ap[c];
      oij[c] = s + d;
      if (i != j) oji[c] = s - d;
    }
  }
}
}  // namespace b200cc

extern "C" int b200cc_pair_rows_unpack(const double* S, const double* A, b200cc_i64 lds, int no, b200cc_i64 ncols,
                                       double* out, b200cc_i64 ldo, void* stream) {
  if (no <= 0 || ncols < 0) { set_error("b200cc_pair_rows_unpack: bad size"); return 1; }
  if (ncols == 0) return 0;
  const i64 np = pair_index(no, 0);
  i64 gx = (ncols + 255) / 256;
  if (gx > 64) gx = 64;
  dim3 grid((unsigned)gx, (unsigned)(np < 65535 ? np : 65535));
  pair_rows_unpack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(S, A, lds, no, ncols, out, ldo);
  return check_launch("pair_rows_unpack_kernel");
}

extern "C" int b200cc_ladder_unpack(const double* S, const double* A, b200cc_i64 lds, int no, int nv, int tri, int a0,
                                    int a1, double alpha, double* r2, void* stream) {
  if (no <= 0 || nv <= 0 || a0 < 0 || a1 > nv || a1 < a0) { set_error("b200cc_ladder_unpack: bad range"); return 1; }
  if (a1 == a0) return 0;
  if (lds < pair_index(a1, 0) - pair_index(a0, 0)) { set_error("b200cc_ladder_unpack: lds too small"); return 1; }
  const int t_lo = a0 / 32, t_hi = (a1 - 1) / 32;
  const i64 nw = pair_index(t_hi + 1, 0) - pair_index(t_lo, 0);
  const i64 M = tri ? pair_index(no, 0) : (i64)no * no;
  dim3 grid((unsigned)nw, (unsigned)(M < 65535 ? M : 65535)), block(32, 8);
  ladder_unpack_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(S, A, lds, no, nv, tri ? 1 : 0, a0, a1, alpha, r2);
  return check_launch("ladder_unpack_kernel");
}
