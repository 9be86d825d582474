// triples.cu -- the Lee-Rendell (T) epilogue, fused.
//
// The six two-segment GEMMs of b200cc_dgemm leave, per (i>=j>=k) triple, Q1..Q6 of shape (nv,nv,nv)
// with   W[a,b,c] = Q1[a,b,c] + Q2[a,c,b] + Q3[c,a,b] + Q4[c,b,a] + Q5[b,c,a] + Q6[b,a,c]
// (the connected numerator of cctriples.py:50-62; particle and hole terms already summed).
// This kernel reads every Q element exactly once and never materialises W, V, X3, Y3, Z3 or the
// denominator cube in global memory: a CTA owns an 8x8x8 block of the sorted (a>=b>=c) index space; the 36 Q
// blocks that make W at the six permutations of (a,b,c) are summed in shared memory, then each thread adds the
// disconnected part on the fly from t1/t2/<ij|ab>/f_ov (cctriples.py:131-137), applies the
// 1/(1+delta) factor (210-213), forms the Lee-Rendell bracket (215-237) and reduces.
#include <stdlib.h>
#include "common.cuh"

namespace b200cc {

constexpr int TT = 8;  // cube edge

struct TArgs {
  int no, nv, nt;
  int blocked;   // Q stored as contiguous 8x8x8 cubes (written so by the TMA GEMM epilogue), else plain (v,v,v)
  const int* ijk;
  const double* Q;
  const double *t1, *t2, *oovv, *fov, *eo, *ev;
  i64 ldf;
};

// element (x,y,z) of one Q array: plain row-major (v,v,v), or 8x8x8-cube-blocked with nc8 = ceil(v/8) cubes per edge
__device__ __forceinline__ i64 qoff(int blocked, int nv, int x, int y, int z) {
  if (!blocked) return ((i64)x * nv + y) * nv + z;
  const int nc8 = (nv + 7) >> 3;
  return ((((i64)(x >> 3) * nc8 + (y >> 3)) * nc8 + (z >> 3)) << 9) + ((x & 7) << 6) + ((y & 7) << 3) + (z & 7);
}
__host__ __device__ __forceinline__ i64 qsize(int blocked, int nv) {
  if (!blocked) return (i64)nv * nv * nv;
  const i64 nc8 = (nv + 7) >> 3;
  return nc8 * nc8 * nc8 * 512;
}

__device__ __forceinline__ double wsum(const double* __restrict__ Q, i64 v3, int blocked, int nv, int x, int y, int z) {
  return Q[qoff(blocked, nv, x, y, z)] + Q[v3 + qoff(blocked, nv, x, z, y)] + Q[2 * v3 + qoff(blocked, nv, z, x, y)] +
         Q[3 * v3 + qoff(blocked, nv, z, y, x)] + Q[4 * v3 + qoff(blocked, nv, y, z, x)] +
         Q[5 * v3 + qoff(blocked, nv, y, x, z)];
}

struct Disc {  // row pointers for the disconnected part of one (i,j,k)
  const double *Kij, *Kik, *Kjk, *Tij, *Tik, *Tjk, *t1i, *t1j, *t1k, *fi, *fj, *fk;
  int nv;
  __device__ __forceinline__ double operator()(int x, int y, int z) const {
    return Kij[x * nv + y] * t1k[z] + Kik[x * nv + z] * t1j[y] + Kjk[y * nv + z] * t1i[x] +
           Tij[x * nv + y] * fk[z] + Tik[x * nv + z] * fj[y] + Tjk[y * nv + z] * fi[x];
  }
};

__device__ __forceinline__ Disc make_disc(const TArgs& p, int i, int j, int k) {
  const i64 vv = (i64)p.nv * p.nv;
  Disc d;
  d.Kij = p.oovv + ((i64)i * p.no + j) * vv; d.Kik = p.oovv + ((i64)i * p.no + k) * vv;
  d.Kjk = p.oovv + ((i64)j * p.no + k) * vv;
  d.Tij = p.t2 + ((i64)i * p.no + j) * vv; d.Tik = p.t2 + ((i64)i * p.no + k) * vv;
  d.Tjk = p.t2 + ((i64)j * p.no + k) * vv;
  d.t1i = p.t1 + (i64)i * p.nv; d.t1j = p.t1 + (i64)j * p.nv; d.t1k = p.t1 + (i64)k * p.nv;
  d.fi = p.fov + (i64)i * p.ldf; d.fj = p.fov + (i64)j * p.ldf; d.fk = p.fov + (i64)k * p.ldf;
  d.nv = p.nv;
  return d;
}

// grid = (sorted 8-cube triples, ntrip); block = 512, two CTAs per SM.
// Phase 1: the 36 (Q_n, permutation) blocks that make W at the six permutations of this cube are read with
// COALESCED 64-byte runs (thread = one element of the source block, fastest index = the block's own last
// dimension) and scatter-added into six 8x8x8 shared tiles W_P[la][lb][lc]; for a fixed n the six targets
// are distinct tiles and thread -> slot is a bijection, so only one __syncthreads per n is needed.
// Phase 2: thread (la,lb,lc) combines its six W values with the on-the-fly disconnected part.
// (Measured on B200: 2.66 TB/s of Q traffic vs 1.93 TB/s for direct per-thread gathers; issuing all 36
//  loads up front at one CTA per SM was slower, 2.04 TB/s.)
constexpr int SA = 73, SB = 9, WTILE = 8 * SA;   // padded tile pitch (doubles)

template <bool BLOCKED, bool HOIST>
__global__ void __launch_bounds__(512, 2) t_energy_kernel(const TArgs p, double* scratch) {
  __shared__ double Wsm[6][WTILE];
  __shared__ double red[16];
  // decode blockIdx.x -> (TA >= TB >= TC)
  int rem = blockIdx.x, TA = 0;
  while ((TA + 1) * (TA + 2) * (TA + 3) / 6 <= rem) ++TA;
  rem -= TA * (TA + 1) * (TA + 2) / 6;
  int TB = 0;
  while ((TB + 1) * (TB + 2) / 2 <= rem) ++TB;
  const int TC = rem - TB * (TB + 1) / 2;
  const int trip = blockIdx.y;
  const int i = p.ijk[3 * trip], j = p.ijk[3 * trip + 1], k = p.ijk[3 * trip + 2];
  const int nv = p.nv;
  const i64 v3 = qsize(p.blocked, nv);
  const double* Q = p.Q + (i64)trip * 6 * v3;
  const int T[3] = {TA * TT, TB * TT, TC * TT};
  const int u[3] = {(int)(threadIdx.x >> 6), (int)((threadIdx.x >> 3) & 7), (int)(threadIdx.x & 7)};

  // P (target permutation of (a,b,c)) and pi_n (index order of Q_n), as position tables
  constexpr int PERM[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  constexpr int PI[6][3] = {{0, 1, 2}, {0, 2, 1}, {2, 0, 1}, {2, 1, 0}, {1, 2, 0}, {1, 0, 2}};
  // cube-blocked Q: a source block is one contiguous 4 KB run (thread t reads element t); plain Q: 64-byte runs
  const int nc8 = (nv + 7) >> 3;
  // HOIST: everything per-(n,P) that does not depend on the thread (block origin, bounds) is CTA-uniform, the thread's
  // offset inside a source block is the same for all 36 blocks, and the shared slot only depends on which of the six
  // axis maps rho applies.  Loads stay PREDICATED (bitwise conditions, no short-circuit branches).
  const i64 tpart = BLOCKED ? (i64)threadIdx.x : ((i64)u[0] * nv + u[1]) * nv + u[2];
  int dsto[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    int l[3];
    l[PERM[r][0]] = u[0]; l[PERM[r][1]] = u[1]; l[PERM[r][2]] = u[2];
    dsto[r] = l[0] * SA + l[1] * SB + l[2];
  }
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    const double* Qn = Q + (i64)n * v3;
#pragma unroll
    for (int P = 0; P < 6; ++P) {
      // Q_n coordinate k of W[P(a,b,c)] is cube axis rho_k = PERM[P][PI[n][k]]
      const int r0 = PERM[P][PI[n][0]], r1 = PERM[P][PI[n][1]], r2 = PERM[P][PI[n][2]];
      double val = 0.0;
      double* dst;
      if (HOIST) {
        const i64 origin = BLOCKED ? ((((i64)(T[r0] >> 3) * nc8 + (T[r1] >> 3)) * nc8 + (T[r2] >> 3)) << 9)
                                   : ((i64)T[r0] * nv + T[r1]) * nv + T[r2];
        const bool ok = (u[0] < nv - T[r0]) & (u[1] < nv - T[r1]) & (u[2] < nv - T[r2]);
        if (ok) val = __ldg(Qn + origin + tpart);
        dst = &Wsm[P][dsto[r0 * 2 + (r1 > r2 ? 1 : 0)]];
      } else {
        const int x = T[r0] + u[0], y = T[r1] + u[1], z = T[r2] + u[2];
        // keep this a single PREDICATED load: with a branch per load (e.g. an `interior ||` short-circuit) the compiler
        // wraps each of the 36 loads in BSSY/BSYNC and they no longer overlap -- measured 1.9x slower (1.4 vs 2.7 TB/s)
        const i64 off = BLOCKED ? (((((i64)(T[r0] >> 3) * nc8 + (T[r1] >> 3)) * nc8 + (T[r2] >> 3)) << 9) + threadIdx.x)
                                : (((i64)x * nv + y) * nv + z);
        if (x < nv && y < nv && z < nv) val = __ldg(Qn + off);
        // cube-local coordinates of this element: l[rho_k] = u_k
        int l[3];
        l[r0] = u[0]; l[r1] = u[1]; l[r2] = u[2];
        dst = &Wsm[P][l[0] * SA + l[1] * SB + l[2]];
      }
      if (n == 0) *dst = val;
      else *dst += val;
    }
    __syncthreads();
  }

  const int la = u[0], lb = u[1], lc = u[2];
  const int a = T[0] + la, b = T[1] + lb, c = T[2] + lc;
  double e = 0.0;
  if (a < nv && b <= a && c <= b) {
    const int s = la * SA + lb * SB + lc;
    const Disc D = make_disc(p, i, j, k);
    const double sc = 1.0 / (1.0 + (a == b ? 1.0 : 0.0) + (a == c ? 1.0 : 0.0) + (b == c ? 1.0 : 0.0));
    const double Wabc = Wsm[0][s], Wacb = Wsm[1][s], Wbac = Wsm[2][s];
    const double Wbca = Wsm[3][s], Wcab = Wsm[4][s], Wcba = Wsm[5][s];
    const double Vabc = (Wabc + D(a, b, c)) * sc, Vacb = (Wacb + D(a, c, b)) * sc;
    const double Vbac = (Wbac + D(b, a, c)) * sc, Vbca = (Wbca + D(b, c, a)) * sc;
    const double Vcab = (Wcab + D(c, a, b)) * sc, Vcba = (Wcba + D(c, b, a)) * sc;
    const double X = Wabc * Vabc + Wacb * Vacb + Wbac * Vbac + Wbca * Vbca + Wcab * Vcab + Wcba * Vcba;
    const double Y = Vabc + Vbca + Vcab, Z = Vacb + Vbac + Vcba;
    const double Wc = Wabc + Wbca + Wcab, Wo = Wacb + Wbac + Wcba;
    const double den = p.eo[i] + p.eo[j] + p.eo[k] - p.ev[a] - p.ev[b] - p.ev[c];
    const double occ = 2.0 - ((i == j ? 1.0 : 0.0) + (i == k ? 1.0 : 0.0) + (j == k ? 1.0 : 0.0));
    e = ((Y - 2.0 * Z) * Wc + (Z - 2.0 * Y) * Wo + 3.0 * X) * occ / den;
  }
  // block reduce over 16 warps
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x < 32) {
    double s2 = threadIdx.x < 16 ? red[threadIdx.x] : 0.0;
    s2 = warp_sum(s2);
    if (threadIdx.x == 0) scratch[(i64)trip * gridDim.x + blockIdx.x] = s2;
  }
}

__global__ void __launch_bounds__(256) t3_assemble_kernel(const TArgs p, int i, int j, int k, int with_denom,
                                                          double* w3, double* v3o) {
  const int nv = p.nv;
  const i64 v3 = (i64)nv * nv * nv;
  const i64 qs = qsize(p.blocked, nv);
  const Disc D = make_disc(p, i, j, k);
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < v3; e += (i64)gridDim.x * blockDim.x) {
    const int c = (int)(e % nv);
    const int b = (int)((e / nv) % nv);
    const int a = (int)(e / ((i64)nv * nv));
    double w = wsum(p.Q, qs, p.blocked, nv, a, b, c);   // t3c_ijk numerator
    double v = D(a, b, c);                   // t3d_ijk numerator
    if (with_denom) {
      const double den = p.eo[i] + p.eo[j] + p.eo[k] - p.ev[a] - p.ev[b] - p.ev[c];
      v /= den;
      w /= den;
    }
    w3[e] = w;
    if (v3o) v3o[e] = v;
  }
}

// ---- (T) densities (cctriples.py:1063-1157) ---------------------------------------------------------------------
// Step 1: connected t3 WITH denominators of a batch of triples, M3[t][a,b,c] = (Q1[a,b,c] + Q2[a,c,b] + Q3[c,a,b] +
// Q4[c,b,a] + Q5[b,c,a] + Q6[b,a,c]) / D.  One CTA per 8x8x8 cube of the FULL (a,b,c) space; the six source blocks are
// read with coalesced 64-byte runs (thread = one element of the source block, fastest index = the block's own last
// dimension) and transposed through shared memory, so every Q element is read exactly once and M3 is written coalesced.
__global__ void __launch_bounds__(512, 2) t3_connected_kernel(int no, int nv, int nt, const int* __restrict__ ijk,
                                                              const double* __restrict__ Q,
                                                              const double* __restrict__ eo,
                                                              const double* __restrict__ ev, double* __restrict__ M3) {
  __shared__ double Wsm[WTILE];
  int rem = blockIdx.x;
  const int TC = rem % nt; rem /= nt;
  const int TB = rem % nt;
  const int TA = rem / nt;
  const int trip = blockIdx.y;
  const int i = ijk[3 * trip], j = ijk[3 * trip + 1], k = ijk[3 * trip + 2];
  const i64 v3 = (i64)nv * nv * nv;
  const double* Qt = Q + (i64)trip * 6 * v3;
  const int T[3] = {TA * TT, TB * TT, TC * TT};
  const int u[3] = {(int)(threadIdx.x >> 6), (int)((threadIdx.x >> 3) & 7), (int)(threadIdx.x & 7)};
  constexpr int PI[6][3] = {{0, 1, 2}, {0, 2, 1}, {2, 0, 1}, {2, 1, 0}, {1, 2, 0}, {1, 0, 2}};
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    const int r0 = PI[n][0], r1 = PI[n][1], r2 = PI[n][2];
    const int x = T[r0] + u[0], y = T[r1] + u[1], z = T[r2] + u[2];
    double val = 0.0;
    if (x < nv && y < nv && z < nv) val = __ldg(Qt + (i64)n * v3 + ((i64)x * nv + y) * nv + z);
    int l[3];
    l[r0] = u[0]; l[r1] = u[1]; l[r2] = u[2];
    double* dst = &Wsm[l[0] * SA + l[1] * SB + l[2]];
    if (n == 0) *dst = val;
    else *dst += val;
    __syncthreads();
  }
  const int a = T[0] + u[0], b = T[1] + u[1], c = T[2] + u[2];
  if (a < nv && b < nv && c < nv) {
    const double den = eo[i] + eo[j] + eo[k] - ev[a] - ev[b] - ev[c];
    M3[(i64)trip * v3 + ((i64)a * nv + b) * nv + c] = Wsm[u[0] * SA + u[1] * SB + u[2]] / den;
  }
}

// Step 2: everything of the (i,j,k) loop body of t3_density that is not a GEMM, for fixed (i,j) and a run of k.
// A CTA owns the 8x8 tile (TA,TB) of (a,b) and sweeps all k of the run and all c cubes: per cube the six permuted
// blocks of M3 are staged in shared memory (coalesced), the disconnected t3 N3 is formed on the fly at the six
// permutations, and the thread at (a,b,c) produces
//   W2 = 2 sym(M3) + sym(N3)   and   P = 2 M3 - M3[acb] - M3[cba]         (GEMM operands, written in TWO layouts each:
//                                        [(a,b)][k][c] for the contractions over (k,c), [a][k][(b,c)] for those over (k,b,c))
// and keeps in registers the running sums of the matrix-vector shaped terms (lines 1128, 1133-1137, 1141, 1146):
//   Goovv[i,j,a,b] += 4 t1[k,c] Z3   X2[i,j,a,b] += (M3 - M3[cba]) f[k,c]                       (sum over k, c)
//   dvv[a] += 1/2 M3 (X3+Y3)   Dov[i,a] += (M3 - M3[cba]) (4 t2[j,k,b,c] - 2 t2[j,k,c,b])
//   S1[i,a] += 2 (M3 - M3[bac]) (2<jk|bc> - <jk|cb>)                                             (sum over k, b, c)
// (a,b) sums are written by their owner CTA (no atomics); per-a sums go to scratch[q][a][TB] and a second-stage
// reduction, so the result is deterministic.
struct T3dArgs {
  int no, nv, nt, i, j, k0, nk;
  const double* M3;
  const double *t1, *t2, *oovv, *fov, *eo, *ev;
  i64 ldf;
  double *W2ab, *W2n, *Pab, *Pn, *Gij, *Xij, *scratch;
};

__global__ void __launch_bounds__(512, 1) t3_density_forms_kernel(const T3dArgs p) {
  __shared__ double Wsm[6][WTILE];
  __shared__ double red[16][3];
  const int nv = p.nv, nt = p.nt, no = p.no;
  const int TA = blockIdx.y, TB = blockIdx.x;
  const int la = (int)(threadIdx.x >> 6), lb = (int)((threadIdx.x >> 3) & 7), lc = (int)(threadIdx.x & 7);
  const int u[3] = {la, lb, lc};
  const int a = TA * TT + la, b = TB * TT + lb;
  const i64 vv = (i64)nv * nv, v3 = vv * nv;
  constexpr int PERM[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  int dsto[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    int l[3];
    l[PERM[r][0]] = u[0]; l[PERM[r][1]] = u[1]; l[PERM[r][2]] = u[2];
    dsto[r] = l[0] * SA + l[1] * SB + l[2];
  }
  const int s = la * SA + lb * SB + lc;
  TArgs q;
  q.no = no; q.nv = nv; q.t1 = p.t1; q.t2 = p.t2; q.oovv = p.oovv; q.fov = p.fov; q.ldf = p.ldf;
  double accG = 0.0, accX = 0.0, accD = 0.0, accO = 0.0, accS = 0.0;
  const bool ab_ok = (a < nv) & (b < nv);
  const double evab = ab_ok ? p.ev[a] + p.ev[b] : 0.0;
  for (int kk = 0; kk < p.nk; ++kk) {
    const int k = p.k0 + kk;
    const Disc D = make_disc(q, p.i, p.j, k);
    const double eijk = p.eo[p.i] + p.eo[p.j] + p.eo[k];
    const double* Mk = p.M3 + (i64)kk * v3;
    const double* Tjk = p.t2 + ((i64)p.j * no + k) * vv;
    const double* Kjk = p.oovv + ((i64)p.j * no + k) * vv;
    const double* t1k = p.t1 + (i64)k * nv;
    const double* fk = p.fov + (i64)k * p.ldf;
    for (int TC = 0; TC < nt; ++TC) {
      const int T[3] = {TA * TT, TB * TT, TC * TT};
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const int x = T[PERM[r][0]] + u[0], y = T[PERM[r][1]] + u[1], z = T[PERM[r][2]] + u[2];
        double val = 0.0;
        if ((x < nv) & (y < nv) & (z < nv)) val = __ldg(Mk + ((i64)x * nv + y) * nv + z);
        Wsm[r][dsto[r]] = val;
      }
      __syncthreads();
      const int c = TC * TT + lc;
      if (ab_ok && c < nv) {
        const double m0 = Wsm[0][s], m1 = Wsm[1][s], m2 = Wsm[2][s], m3 = Wsm[3][s], m4 = Wsm[4][s], m5 = Wsm[5][s];
        const double rden = 1.0 / (eijk - evab - p.ev[c]);
        const double n0 = D(a, b, c), n1 = D(a, c, b), n2 = D(b, a, c), n3 = D(b, c, a), n4 = D(c, a, b), n5 = D(c, b, a);
        const double X3 = 8.0 * m0 - 4.0 * (m1 + m2 + m5) + 2.0 * (m3 + m4);
        const double Y3 = (8.0 * n0 - 4.0 * (n1 + n2 + n5) + 2.0 * (n3 + n4)) * rden;
        const double W2 = 2.0 * X3 + Y3;
        const double P = 2.0 * m0 - m1 - m5;
        const double U = m0 - m5;
        const double Z3 = 2.0 * (m0 - m1) - (m2 - m3);
        const i64 iab = (((i64)a * nv + b) * p.nk + kk) * nv + c;
        const i64 in = (((i64)a * p.nk + kk) * nv + b) * nv + c;
        p.W2ab[iab] = W2; p.W2n[in] = W2;
        p.Pab[iab] = P;   p.Pn[in] = P;
        accG += 4.0 * t1k[c] * Z3;
        accX += U * fk[c];
        accD += 0.5 * m0 * (X3 + Y3);
        accO += U * (4.0 * Tjk[(i64)b * nv + c] - 2.0 * Tjk[(i64)c * nv + b]);
        accS += 2.0 * (m0 - m2) * (2.0 * Kjk[(i64)b * nv + c] - Kjk[(i64)c * nv + b]);
      }
      __syncthreads();
    }
  }
  // (a,b) sums: reduce over the 8 lanes that share (la,lb)
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    accG += __shfl_down_sync(0xffffffffu, accG, o, 8);
    accX += __shfl_down_sync(0xffffffffu, accX, o, 8);
  }
  if (lc == 0 && ab_ok) {
    p.Gij[(i64)a * nv + b] += accG;
    p.Xij[(i64)a * nv + b] += accX;
  }
  // per-a sums: the 64 threads of one la are warps 2la and 2la+1
  accD = warp_sum(accD); accO = warp_sum(accO); accS = warp_sum(accS);
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5][0] = accD; red[threadIdx.x >> 5][1] = accO; red[threadIdx.x >> 5][2] = accS;
  }
  __syncthreads();
  if (threadIdx.x < 24) {
    const int w = threadIdx.x / 3, qq = threadIdx.x % 3;
    const int aa = TA * TT + w;
    if (aa < nv) p.scratch[((i64)qq * nv + aa) * nt + TB] = red[2 * w][qq] + red[2 * w + 1][qq];
  }
}

static int sorted_cubes(int nv) {
  const int nt = (nv + TT - 1) / TT;
  return nt * (nt + 1) * (nt + 2) / 6;
}

}  // namespace b200cc

using namespace b200cc;

extern "C" b200cc_i64 b200cc_t_energy_scratch(int nv, int ntrip) {
  return (b200cc_i64)sorted_cubes(nv) * ntrip;
}

extern "C" b200cc_i64 b200cc_t_q_size(int nv, int blocked) { return qsize(blocked, nv); }

extern "C" int b200cc_t_energy_batch(int no, int nv, int ntrip, const int* ijk, const double* Q, int q_blocked,
                                     const double* t1,
                                     const double* t2, const double* oovv, const double* fov, b200cc_i64 ldf,
                                     const double* eo, const double* ev, double* et_out, int accumulate,
                                     double* scratch, void* stream) {
  if (ntrip <= 0 || nv <= 0) return 0;
  if (ntrip > 65535) { set_error("b200cc_t_energy_batch: ntrip > 65535"); return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TArgs p;
  p.no = no; p.nv = nv; p.nt = (nv + TT - 1) / TT; p.blocked = q_blocked ? 1 : 0;
  p.ijk = ijk; p.Q = Q; p.t1 = t1; p.t2 = t2; p.oovv = oovv; p.fov = fov; p.eo = eo; p.ev = ev; p.ldf = ldf;
  const int ncube = sorted_cubes(nv);
  // hoisted index arithmetic measured 2.47 vs 2.03 TB/s on B200 (profiles/t_probe_r01_hoist.log); B200CC_T_HOIST=0 selects
  // the per-element form
  static const bool hoist = [] { const char* e = getenv("B200CC_T_HOIST"); return !(e && e[0] == '0'); }();
  if (p.blocked) t_energy_kernel<true, false><<<dim3(ncube, ntrip), 512, 0, st>>>(p, scratch);
  else if (hoist) t_energy_kernel<false, true><<<dim3(ncube, ntrip), 512, 0, st>>>(p, scratch);
  else t_energy_kernel<false, false><<<dim3(ncube, ntrip), 512, 0, st>>>(p, scratch);
  if (check_launch("t_energy_kernel")) return 1;
  const i64 nparts = (i64)ncube * ntrip;
  if (nparts > 2147483647LL) { set_error("b200cc_t_energy_batch: too many partials"); return 1; }
  return launch_final_reduce(scratch, (int)nparts, 0, 1, et_out, accumulate, 1.0, st);
}

extern "C" int b200cc_t3_assemble(int no, int nv, int i, int j, int k, const double* Q, int q_blocked,
                                  const double* t1,
                                  const double* t2, const double* oovv, const double* fov, b200cc_i64 ldf,
                                  const double* eo, const double* ev, int with_denom, double* w3_out,
                                  double* v3_out, void* stream) {
  if (nv <= 0) return 0;
  TArgs p;
  p.no = no; p.nv = nv; p.nt = (nv + TT - 1) / TT; p.blocked = q_blocked ? 1 : 0;
  p.ijk = nullptr; p.Q = Q; p.t1 = t1; p.t2 = t2; p.oovv = oovv; p.fov = fov; p.eo = eo; p.ev = ev; p.ldf = ldf;
  const i64 v3 = (i64)nv * nv * nv;
  i64 blocks = (v3 + 255) / 256;
  const int cap = sm_count() * 16;
  if (blocks > cap) blocks = cap;
  t3_assemble_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, i, j, k, with_denom,
                                                                                    w3_out, v3_out);
  return check_launch("t3_assemble_kernel");
}

extern "C" int b200cc_t3_connected_batch(int no, int nv, int ntrip, const int* ijk, const double* Q, const double* eo,
                                         const double* ev, double* m3_out, void* stream) {
  if (ntrip <= 0 || nv <= 0) return 0;
  if (ntrip > 65535) { set_error("b200cc_t3_connected_batch: ntrip > 65535"); return 1; }
  const int nt = (nv + TT - 1) / TT;
  t3_connected_kernel<<<dim3((unsigned)(nt * nt * nt), ntrip), 512, 0, static_cast<cudaStream_t>(stream)>>>(
      no, nv, nt, ijk, Q, eo, ev, m3_out);
  return check_launch("t3_connected_kernel");
}

extern "C" b200cc_i64 b200cc_t3_density_scratch(int nv) { return (b200cc_i64)3 * nv * ((nv + TT - 1) / TT); }

extern "C" int b200cc_t3_density_forms(const b200cc_t3d_desc* d, void* stream) {
  if (d->nk <= 0 || d->nv <= 0) return 0;
  if (d->i < 0 || d->i >= d->no || d->j < 0 || d->j >= d->no || d->k0 < 0 || d->k0 + d->nk > d->no) {
    set_error("b200cc_t3_density_forms: occupied indices out of range");
    return 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  T3dArgs p;
  p.no = d->no; p.nv = d->nv; p.nt = (d->nv + TT - 1) / TT; p.i = d->i; p.j = d->j; p.k0 = d->k0; p.nk = d->nk;
  p.M3 = d->M3; p.t1 = d->t1; p.t2 = d->t2; p.oovv = d->oovv; p.fov = d->fov; p.eo = d->eo; p.ev = d->ev;
  p.ldf = d->ldf; p.W2ab = d->W2ab; p.W2n = d->W2n; p.Pab = d->Pab; p.Pn = d->Pn; p.Gij = d->Gij; p.Xij = d->Xij;
  p.scratch = d->scratch;
  t3_density_forms_kernel<<<dim3(p.nt, p.nt), 512, 0, st>>>(p);
  if (check_launch("t3_density_forms_kernel")) return 1;
  const i64 stride = (i64)d->nv * p.nt;
  if (launch_final_reduce(d->scratch, p.nt, p.nt, d->nv, d->dvv, 1, 1.0, st)) return 1;
  if (launch_final_reduce(d->scratch + stride, p.nt, p.nt, d->nv, d->Dov, 1, 1.0, st)) return 1;
  return launch_final_reduce(d->scratch + 2 * stride, p.nt, p.nt, d->nv, d->S1, 1, 1.0, st);
}
