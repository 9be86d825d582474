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
  int paired;    // 3 arrays per triple instead of 6: R1[x,y,z] = Q1[x,y,z] + Q6[y,x,z], R2 = Q2 + Q3^T, R3 = Q4 + Q5^T (the
                 // GEMM summed each pair in its accumulators), W[a,b,c] = R1[a,b,c] + R2[a,c,b] + R3[c,b,a]
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

__device__ __forceinline__ double wsum(const double* __restrict__ Q, i64 v3, int blocked, int nv, int x, int y, int z,
                                       int paired = 0) {
  if (paired)
    return Q[qoff(blocked, nv, x, y, z)] + Q[v3 + qoff(blocked, nv, x, z, y)] + Q[2 * v3 + qoff(blocked, nv, z, y, x)];
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
// array n of NQ (6 plain, 3 paired) has the index order of Q_{QSRC(n)}: paired R1, R2, R3 are laid out like Q1, Q2, Q4
template <int NQ>
__host__ __device__ constexpr int qsrc(int n) { return NQ == 6 ? n : (n == 2 ? 3 : n); }

template <bool BLOCKED, bool HOIST, int NQ>
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
  const double* Q = p.Q + (i64)trip * NQ * v3;
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
  for (int n = 0; n < NQ; ++n) {
    const double* Qn = Q + (i64)n * v3;
#pragma unroll
    for (int P = 0; P < 6; ++P) {
      // Q_n coordinate k of W[P(a,b,c)] is cube axis rho_k = PERM[P][PI[n][k]]
      const int r0 = PERM[P][PI[qsrc<NQ>(n)][0]], r1 = PERM[P][PI[qsrc<NQ>(n)][1]], r2 = PERM[P][PI[qsrc<NQ>(n)][2]];
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
    double w = wsum(p.Q, qs, p.blocked, nv, a, b, c, p.paired);   // t3c_ijk numerator
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

// Conflict-free 8x8x8 shared tile: element (l0,l1,l2) lives at ((l0,l1,l2) as 9 bits) ^ h(l0,l1) with
//   h = 9 (l0 & 1) ^ (l0 & 6) ^ (l1 & 6),
// a GF(2)-linear swizzle of the low four address bits (= the 64-bit bank).  For EVERY assignment of the thread
// coordinates (la,lb,lc) to (l0,l1,l2) the 16 lanes of a half-warp (lc = 0..7, two lb) hit 16 distinct banks, so the
// six permuted scatter-stores and the natural-order loads are all single-wavefront (the padded 73/9 pitches of the
// energy kernel are 2-way conflicted on every pattern; ncu: 1.5 extra wavefronts per shared instruction).
__device__ __forceinline__ int swz(int l0, int l1, int l2) {
  return ((l0 << 6) | (l1 << 3) | l2) ^ (((l0 & 1) * 9) ^ (l0 & 6) ^ (l1 & 6));
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = pred ? 8 : 0;            // src-size 0: nothing is read, the 8 bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// cp.async-staged variant of the (T) energy kernel (plain Q layout).  Two things bound the classic kernel above
// (2.5 TB/s): its six load / barrier rounds keep only ~50 KB per SM in flight, and -- mostly -- phase 2 evaluates the
// disconnected t3 at the six permutations with 72 scattered global loads per thread (54 distinct), twice the number of
// Q loads.  Here
//  * the 36 (Q_n, permutation) blocks of a cube are copied asynchronously, 18 at a time (72 KB), straight into
//    conflict-free swizzled slots (element l of target permutation P from Q_n lands in raw[n%3][P][swz(l)]); each thread
//    then adds its six W values out of shared memory;
//  * the 36 8x8 tiles of <ij|..>, <ik|..>, <jk|..>, t2[ij], t2[ik], t2[jk] over the ordered pairs of (A,B,C), and the
//    t1 / f_ov / eps_v segments of the three ranges, ride in the first copy group (4.5 + 0.33 copies per thread), so the
//    disconnected part is formed from shared memory only.
// Two CTAs per SM overlap one CTA's copies with the other's arithmetic.
constexpr int TE_ROUND = 18 * 512;          // doubles staged per round
constexpr int TE_TP = 9;                    // tile pitch
constexpr int TE_TILES = 36 * 8 * TE_TP;    // [mat 6][ordered pair 6][8][pitch]
constexpr int TE_VECS = 7 * 24;             // t1[i], t1[j], t1[k], f[i], f[j], f[k], eps_v  x  ranges A, B, C  x  8
constexpr int TE_DOUBLES = TE_ROUND + TE_TILES + TE_VECS;

// ordered pair (X,Y), X != Y, of the axes {0:a, 1:b, 2:c}  ->  0..5
__host__ __device__ constexpr int te_pair(int X, int Y) { return X == 0 ? (Y == 1 ? 0 : 2) : (X == 1 ? (Y == 0 ? 1 : 4) : (Y == 0 ? 3 : 5)); }

// disconnected t3 numerator (cctriples.py:131-137) at the permutation (x,y,z) = (axis X, axis Y, axis Z) of this
// thread's (a,b,c), from the staged tiles: mat 0..2 = <ij|, <ik|, <jk|; 3..5 = t2[ij], t2[ik], t2[jk]; vec 0..2 = t1[i],
// t1[j], t1[k]; 3..5 = f[i], f[j], f[k]
template <int X, int Y, int Z>
__device__ __forceinline__ double te_disc(const double* tiles, const double* vecs, const int (&l)[3]) {
  auto tl = [&](int mat, int P, int Q) { return tiles[((mat * 6 + te_pair(P, Q)) * 8 + l[P]) * TE_TP + l[Q]]; };
  auto vc = [&](int vec, int P) { return vecs[(vec * 3 + P) * 8 + l[P]]; };
  return tl(0, X, Y) * vc(2, Z) + tl(1, X, Z) * vc(1, Y) + tl(2, Y, Z) * vc(0, X) +
         tl(3, X, Y) * vc(5, Z) + tl(4, X, Z) * vc(4, Y) + tl(5, Y, Z) * vc(3, X);
}

template <int NQ>
__global__ void __launch_bounds__(512, 2) t_energy_cp_kernel(const TArgs p, double* scratch) {
  extern __shared__ double raw[];    // [3][6][512] | tiles | vecs
  __shared__ double red[16];
  double* tiles = raw + TE_ROUND;
  double* vecs = tiles + TE_TILES;
  int rem = blockIdx.x, TA = 0;
  while ((TA + 1) * (TA + 2) * (TA + 3) / 6 <= rem) ++TA;
  rem -= TA * (TA + 1) * (TA + 2) / 6;
  int TB = 0;
  while ((TB + 1) * (TB + 2) / 2 <= rem) ++TB;
  const int TC = rem - TB * (TB + 1) / 2;
  const int trip = blockIdx.y;
  const int i = p.ijk[3 * trip], j = p.ijk[3 * trip + 1], k = p.ijk[3 * trip + 2];
  const int nv = p.nv, no = p.no;
  const i64 vv = (i64)nv * nv, v3 = vv * nv;
  const double* Q = p.Q + (i64)trip * NQ * v3;
  const int T[3] = {TA * TT, TB * TT, TC * TT};
  const int u[3] = {(int)(threadIdx.x >> 6), (int)((threadIdx.x >> 3) & 7), (int)(threadIdx.x & 7)};
  constexpr int PERM[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  constexpr int PI[6][3] = {{0, 1, 2}, {0, 2, 1}, {2, 0, 1}, {2, 1, 0}, {1, 2, 0}, {1, 0, 2}};
  const i64 tpart = ((i64)u[0] * nv + u[1]) * nv + u[2];
  // shared slot of this thread's element for each of the six axis maps rho (index r0*2 + (r1 > r2)), see classic kernel
  int dsto[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    int l[3];
    l[PERM[r][0]] = u[0]; l[PERM[r][1]] = u[1]; l[PERM[r][2]] = u[2];
    dsto[r] = swz(l[0], l[1], l[2]);
  }
  const int s = swz(u[0], u[1], u[2]);
  // ---- tiles and vectors of the disconnected part (first copy group) --------------------------------------------
  {
    const i64 pij = ((i64)i * no + j) * vv, pik = ((i64)i * no + k) * vv, pjk = ((i64)j * no + k) * vv;
    const int row = u[1], col = u[2];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int t = (int)(threadIdx.x >> 6) + 8 * q;          // tile 0..35 (q = 4: only tiles 32..35)
      if (t < 36) {
        const int mat = t / 6, pr = t - mat * 6;
        // ordered pair pr -> (X,Y):  0 (a,b)  1 (b,a)  2 (a,c)  3 (c,a)  4 (b,c)  5 (c,b)
        const int X = pr == 0 || pr == 2 ? 0 : (pr == 1 || pr == 4 ? 1 : 2);
        const int Y = pr == 1 || pr == 3 ? 0 : (pr == 0 || pr == 5 ? 1 : 2);
        const int m3 = mat % 3;
        const double* base = (mat < 3 ? p.oovv : p.t2) + (m3 == 0 ? pij : (m3 == 1 ? pik : pjk));
        const int x = T[X] + row, y = T[Y] + col;
        const bool ok = (x < nv) & (y < nv);
        cp_async8(tiles + (t * 8 + row) * TE_TP + col, base + (ok ? (i64)x * nv + y : 0), ok);
      }
    }
    if (threadIdx.x < TE_VECS) {
      const int vec = (int)threadIdx.x / 24, rg = ((int)threadIdx.x % 24) >> 3, el = (int)threadIdx.x & 7;
      const int occ = vec % 3 == 0 ? i : (vec % 3 == 1 ? j : k);
      const double* src = vec == 6 ? p.ev : (vec < 3 ? p.t1 + (i64)occ * nv : p.fov + (i64)occ * p.ldf);
      const int x = T[rg] + el;
      const bool ok = x < nv;
      cp_async8(vecs + threadIdx.x, src + (ok ? x : 0), ok);
    }
  }
  double W[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int half = 0; half < NQ / 3; ++half) {      // 18 source blocks per round: two rounds plain, one paired
#pragma unroll
    for (int nn = 0; nn < 3; ++nn) {
      const int n = half * 3 + nn;
      const double* Qn = Q + (i64)n * v3;
#pragma unroll
      for (int P = 0; P < 6; ++P) {
        const int r0 = PERM[P][PI[qsrc<NQ>(n)][0]], r1 = PERM[P][PI[qsrc<NQ>(n)][1]], r2 = PERM[P][PI[qsrc<NQ>(n)][2]];
        const i64 origin = ((i64)T[r0] * nv + T[r1]) * nv + T[r2];
        const bool ok = (u[0] < nv - T[r0]) & (u[1] < nv - T[r1]) & (u[2] < nv - T[r2]);
        cp_async8(raw + (nn * 6 + P) * 512 + dsto[r0 * 2 + (r1 > r2 ? 1 : 0)], Qn + (ok ? origin + tpart : 0), ok);
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll
    for (int P = 0; P < 6; ++P) W[P] += raw[P * 512 + s] + raw[(6 + P) * 512 + s] + raw[(12 + P) * 512 + s];
    if (half + 1 < NQ / 3) __syncthreads();   // everyone has read this round before the next overwrites it
  }

  const int l[3] = {u[0], u[1], u[2]};
  const int a = T[0] + l[0], b = T[1] + l[1], c = T[2] + l[2];
  double e = 0.0;
  if (a < nv && b <= a && c <= b) {
    const double sc = 1.0 / (1.0 + (a == b ? 1.0 : 0.0) + (a == c ? 1.0 : 0.0) + (b == c ? 1.0 : 0.0));
    const double Wabc = W[0], Wacb = W[1], Wbac = W[2], Wbca = W[3], Wcab = W[4], Wcba = W[5];
    const double Vabc = (Wabc + te_disc<0, 1, 2>(tiles, vecs, l)) * sc, Vacb = (Wacb + te_disc<0, 2, 1>(tiles, vecs, l)) * sc;
    const double Vbac = (Wbac + te_disc<1, 0, 2>(tiles, vecs, l)) * sc, Vbca = (Wbca + te_disc<1, 2, 0>(tiles, vecs, l)) * sc;
    const double Vcab = (Wcab + te_disc<2, 0, 1>(tiles, vecs, l)) * sc, Vcba = (Wcba + te_disc<2, 1, 0>(tiles, vecs, l)) * sc;
    const double X = Wabc * Vabc + Wacb * Vacb + Wbac * Vbac + Wbca * Vbca + Wcab * Vcab + Wcba * Vcba;
    const double Y = Vabc + Vbca + Vcab, Z = Vacb + Vbac + Vcba;
    const double Wc = Wabc + Wbca + Wcab, Wo = Wacb + Wbac + Wcba;
    const double den = p.eo[i] + p.eo[j] + p.eo[k] - vecs[(6 * 3 + 0) * 8 + l[0]] - vecs[(6 * 3 + 1) * 8 + l[1]] -
                       vecs[(6 * 3 + 2) * 8 + l[2]];
    const double occ = 2.0 - ((i == j ? 1.0 : 0.0) + (i == k ? 1.0 : 0.0) + (j == k ? 1.0 : 0.0));
    e = ((Y - 2.0 * Z) * Wc + (Z - 2.0 * Y) * Wo + 3.0 * X) * occ / den;
  }
  e = warp_sum(e);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x < 32) {
    double s2 = threadIdx.x < 16 ? red[threadIdx.x] : 0.0;
    s2 = warp_sum(s2);
    if (threadIdx.x == 0) scratch[(i64)trip * gridDim.x + blockIdx.x] = s2;
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
  double val[6];               // all six loads in flight before the first barrier
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    const int x = T[PI[n][0]] + u[0], y = T[PI[n][1]] + u[1], z = T[PI[n][2]] + u[2];
    val[n] = 0.0;
    if ((x < nv) & (y < nv) & (z < nv)) val[n] = __ldg(Qt + (i64)n * v3 + ((i64)x * nv + y) * nv + z);
  }
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    int l[3];
    l[PI[n][0]] = u[0]; l[PI[n][1]] = u[1]; l[PI[n][2]] = u[2];
    double* dst = &Wsm[l[0] * SA + l[1] * SB + l[2]];
    if (n == 0) *dst = val[n];
    else *dst += val[n];
    __syncthreads();
  }
  const int a = T[0] + u[0], b = T[1] + u[1], c = T[2] + u[2];
  if (a < nv && b < nv && c < nv) {
    const double den = eo[i] + eo[j] + eo[k] - ev[a] - ev[b] - ev[c];
    M3[(i64)trip * v3 + ((i64)a * nv + b) * nv + c] = Wsm[u[0] * SA + u[1] * SB + u[2]] / den;
  }
}

// Step 2: everything of the (i,j,k) loop body of t3_density that is not a GEMM, for fixed (i,j) and a run of k.
// A CTA owns the 8x8 tile (TA,TB) of (a,b) and sweeps all k of the run and all c cubes.  Per cube the six permuted
// blocks of M3 are staged in shared memory (coalesced 64-byte runs) and the thread at (a,b,c) produces
//   W2 = 2 sym(M3) + sym(N3)   and   P = 2 M3 - M3[acb] - M3[cba]         (GEMM operands, written in TWO layouts each:
//                                        [(a,b)][k][c] for the contractions over (k,c), [a][k][(b,c)] for those over (k,b,c))
// and keeps in registers the running sums of the matrix-vector shaped terms (lines 1128, 1133-1137, 1141, 1146):
//   Goovv[i,j,a,b] += 4 t1[k,c] Z3   X2[i,j,a,b] += (M3 - M3[cba]) f[k,c]                       (sum over k, c)
//   dvv[a] += 1/2 M3 (X3+Y3)   Dov[i,a] += (M3 - M3[cba]) (4 t2[j,k,b,c] - 2 t2[j,k,c,b])
//   S1[i,a] += 2 (M3 - M3[bac]) (2<jk|bc> - <jk|cb>)                                             (sum over k, b, c)
// (a,b) sums are written by their owner CTA (no atomics); per-a sums go to scratch[q][a][TB] and a second-stage
// reduction, so the result is deterministic.
//
// The disconnected t3 is never formed per permutation.  With A~[x,y] = 4 A[x,y] - 2 A[y,x] the symmetriser of
// cctriples.py:1125 applied to the three outer-product shapes of t3d_ijk (cctriples.py:131-137) collapses to
//   sym(A[x,y] w[z]) = 2 A~[a,b] w[c] -   A~[a,c] w[b] -   A~[c,b] w[a]
//   sym(A[x,z] w[y]) = 2 A~[a,c] w[b] -   A~[b,c] w[a] -   A~[a,b] w[c]
//   sym(A[y,z] w[x]) = 2 A~[b,c] w[a] -   A~[a,c] w[b] -   A~[b,a] w[c]
// so per cube only twelve 8x8 tiles of K~ = 4<pq|ab> - 2<pq|ba> and T~ = 4 t2 - 2 t2^T (both built once by the host)
// are staged next to the M3 blocks; the (a,b)-only factors live in registers.  K~[j,k][b,c] and T~[j,k][b,c] are also
// exactly the weights of the S1 (x 1/2) and Dov sums.  Loads of the next cube are issued before the arithmetic of the
// current one (register prefetch).  swap_ab: M3 holds the run of the TRANSPOSED pair, M3(j,i,k)[b,a,c] = M3(i,j,k)[a,b,c]
// (cctriples.py:50-62 is symmetric under simultaneous permutation of (i,j,k) and (a,b,c)), so one t3 build serves both.
struct T3dArgs {
  int no, nv, nt, i, j, k0, nk, swap_ab;
  const double* M3;
  const double *t1, *Ts, *Ks, *fov, *eo, *ev;
  i64 ldf;
  double *W2ab, *W2n, *Pab, *Pn, *Gij, *Xij, *scratch;
};

constexpr int T3D_NST = 4;                    // ring depth
constexpr int T3D_STAGE = 6 * 512 + 12 * 64 + 64;  // doubles per stage: M3 blocks, K~/T~ tiles, seven c-vectors

template <bool SWAP>
__global__ void __launch_bounds__(512, 1) t3_density_forms_kernel(const T3dArgs p) {
  // NST-deep ring of stages filled by cp.async (8-byte copies straight into the swizzled slots, zero-fill out of range):
  // stage = six 8x8x8 M3 blocks + twelve 8x8 K~/T~ tiles = 30 KB.  With one CTA per SM and ~2 us of loaded DRAM latency a
  // single register-prefetched stage kept only 32 KB per SM in flight (measured 1.9 TB/s of DRAM traffic); the ring
  // keeps NST-1 stages in flight and needs ONE __syncthreads per cube.
  extern __shared__ double dsm[];
  __shared__ double red[16][3];
  const int nv = p.nv, nt = p.nt, no = p.no;
  // (giving the twin tiles (x,y) / (y,x), which read the same six M3 blocks, adjacent block indices was measured
  //  SLOWER, 59.5 vs 54.4 ms at o=40,v=300: the row-major order already keeps a wave inside a few a-rows)
  const int TA = blockIdx.y, TB = blockIdx.x;
  const int la = (int)(threadIdx.x >> 6), lb = (int)((threadIdx.x >> 3) & 7), lc = (int)(threadIdx.x & 7);
  const int u[3] = {la, lb, lc};
  const int a = TA * TT + la, b = TB * TT + lb;
  const i64 vv = (i64)nv * nv, v3 = vv * nv;
  constexpr int PERM[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  const int TAB[2] = {TA * TT, TB * TT};

  // ---- prefetch state: everything advances by constant pointer steps per c-cube / per k -------------------------
  // M3 block of target permutation r: source coordinates (x,y,z) = (X[PERM[r][0]] + u0, X[PERM[r][1]] + u1, X[PERM[r][2]] + u2)
  // with X = (A origin, B origin, C origin); the C origin moves by 8 per iteration along the source axis holding c.
  unsigned moff[6];          // element offsets into M3 (the host keeps nk * nv^3 < 2^32)
  unsigned mstep[6];
  int dsto[6];
  unsigned abok = 0;         // bit r: the two non-c source coordinates are in range
  int ucp[6];                // thread coordinate that sits on the c axis of block r
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    int l[3];
    l[PERM[r][0]] = u[0]; l[PERM[r][1]] = u[1]; l[PERM[r][2]] = u[2];
    dsto[r] = swz(l[0], l[1], l[2]);
    unsigned off = 0;
    bool ok = true;
    const unsigned st[3] = {(unsigned)vv, (unsigned)nv, 1u};
    mstep[r] = 0; ucp[r] = 0;
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      if (PERM[r][n] == 2) { mstep[r] = TT * st[n]; ucp[r] = u[n]; off += (unsigned)u[n] * st[n]; }
      else { const int x = TAB[PERM[r][n]] + u[n]; ok = ok && (x < nv); off += (unsigned)(ok ? x : 0) * st[n]; }
    }
    moff[r] = off;
    if (ok) abok |= 1u << r;
  }
  // the tile(s) this thread stages: tile t = tid/64 (and 8 + tid/64 for tid < 256), as 8x8 elements (lb, lc)
  //   t % 6:  0 ij[a,c]  1 ij[c,b]  2 ik[a,c]  3 ik[b,c]  4 jk[b,c]  5 jk[a,c];   t / 6: 0 = K~, 1 = T~
  // every tile is stored [row = its a- or b-coordinate][col = c]; for tile 1 (source rows = c) the thread reads the
  // source element (c = C + lc, b = B + lb) -- strided in global memory, conflict-free in shared memory.
  const int tl[2] = {(int)(threadIdx.x >> 6), 8 + (int)(threadIdx.x >> 6)};
  const bool has1 = threadIdx.x < 256;
  const double* tbase[2];
  unsigned toff[2], tstep[2], tkstep[2];     // element offsets into K~ / T~ (no^2 nv^2 < 2^32, checked by the host)
  bool tok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = tl[h] % 6;
    tbase[h] = tl[h] >= 6 ? p.Ts : p.Ks;
    const unsigned base = (unsigned)(m < 2 ? (p.i * no + p.j) : (m < 4 ? (p.i * no + p.k0) : (p.j * no + p.k0))) * (unsigned)vv;
    tkstep[h] = m < 2 ? 0u : (unsigned)vv;
    const int rowo = (m == 3 || m == 4 || m == 1) ? TAB[1] : TAB[0];       // b-rows for bc / cb tiles, else a-rows
    const int x = rowo + lb;
    tok[h] = (x < nv) && (h == 0 || has1);
    const unsigned xs = tok[h] ? (unsigned)x : 0u;
    if (m == 1) { toff[h] = base + (unsigned)lc * nv + xs; tstep[h] = (unsigned)TT * nv; }
    else { toff[h] = base + xs * nv + lc; tstep[h] = TT; }
  }
  int nTC = 0, nkk = 0;      // (k, c-cube) of the NEXT iteration to load
  auto issue_next = [&](int stage) {
    double* Ws = dsm + stage * T3D_STAGE;
    const int c0 = nTC * TT;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const bool ok = ((abok >> r) & 1u) && (c0 + ucp[r] < nv);
      cp_async8(Ws + r * 512 + dsto[r], p.M3 + (ok ? moff[r] : 0u), ok);
      moff[r] += mstep[r];
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 0 || has1) {
        const bool ok = tok[h] && (c0 + lc < nv);
        cp_async8(Ws + 3072 + tl[h] * 64 + lb * 8 + lc, tbase[h] + (ok ? toff[h] : 0u), ok);
      }
      toff[h] += tstep[h];
    }
    if (threadIdx.x < 56) {
      // the c-dependent vectors of this cube: t1[k], t1[j], t1[i], f[k], f[j], f[i], eps_v  (vector q = tid / 8)
      const int q = (int)(threadIdx.x >> 3), k = p.k0 + nkk;
      const double* src = q == 6 ? p.ev
                                 : (q < 3 ? p.t1 + (i64)(q == 0 ? k : (q == 1 ? p.j : p.i)) * nv
                                          : p.fov + (i64)(q == 3 ? k : (q == 4 ? p.j : p.i)) * p.ldf);
      const bool ok = c0 + lc < nv;
      cp_async8(Ws + 3840 + threadIdx.x, src + (ok ? c0 + lc : 0), ok);
    }
    if (++nTC == nt) {
      nTC = 0; ++nkk;
#pragma unroll
      for (int r = 0; r < 6; ++r) moff[r] += (unsigned)v3 - (unsigned)nt * mstep[r];
#pragma unroll
      for (int h = 0; h < 2; ++h) toff[h] += tkstep[h] - (unsigned)nt * tstep[h];
    }
  };

  // ---- (a,b)-only factors ---------------------------------------------------------------------------------------
  const bool ab_ok = (a < nv) & (b < nv);
  const int ac = a < nv ? a : 0, bc = b < nv ? b : 0;
  const i64 pij = ((i64)p.i * no + p.j) * vv;
  const double* t1i = p.t1 + (i64)p.i * nv;
  const double* t1j = p.t1 + (i64)p.j * nv;
  const double* fi = p.fov + (i64)p.i * p.ldf;
  const double* fj = p.fov + (i64)p.j * p.ldf;
  const double t1ia = t1i[ac], t1ib = t1i[bc], t1ja = t1j[ac], t1jb = t1j[bc];
  const double fia = fi[ac], fib = fi[bc], fja = fj[ac], fjb = fj[bc];
  const double Kij_ab = p.Ks[pij + (i64)ac * nv + bc], Tij_ab = p.Ts[pij + (i64)ac * nv + bc];
  const double evab = p.ev[ac] + p.ev[bc];
  const double eij = p.eo[p.i] + p.eo[p.j];
  // M3 of the transposed pair: target permutation r reads the tile of r with its first two source axes exchanged
  const int s = swz(la, lb, lc);
  const int ac_ = la * 8 + lc, bc_ = lb * 8 + lc;
  // output cursors: W2ab/Pab [(a,b)][kk][c], W2n/Pn [a][kk][b][c]
  i64 oab = ((i64)ac * nv + bc) * p.nk * nv + lc;
  i64 on = (i64)ac * p.nk * vv + (i64)bc * nv + lc;

  double accG = 0.0, accX = 0.0, accD = 0.0, accO = 0.0, accS = 0.0;
  double Kik_ab = 0.0, Tik_ab = 0.0, Kjk_ba = 0.0, Tjk_ba = 0.0, t1ka = 0.0, t1kb = 0.0, fka = 0.0, fkb = 0.0, eabk = 0.0;
  const double *t1k = p.t1, *fk = p.fov;
#pragma unroll
  for (int st0 = 0; st0 < T3D_NST - 1; ++st0) {
    if (nkk < p.nk) issue_next(st0);
    cp_async_commit();
  }
  int stage = 0;
  for (int kk = 0; kk < p.nk; ++kk) {
    {
      const int k = p.k0 + kk;
      const i64 pik = ((i64)p.i * no + k) * vv, pjk = ((i64)p.j * no + k) * vv;
      Kik_ab = p.Ks[pik + (i64)ac * nv + bc]; Tik_ab = p.Ts[pik + (i64)ac * nv + bc];
      Kjk_ba = p.Ks[pjk + (i64)bc * nv + ac]; Tjk_ba = p.Ts[pjk + (i64)bc * nv + ac];
      t1k = p.t1 + (i64)k * nv; fk = p.fov + (i64)k * p.ldf;
      t1ka = t1k[ac]; t1kb = t1k[bc]; fka = fk[ac]; fkb = fk[bc];
      eabk = eij + p.eo[k] - evab;
    }
    for (int TC = 0; TC < nt; ++TC) {
      cp_async_wait<T3D_NST - 2>();
      __syncthreads();          // stage `stage` has landed for everyone; everyone is done with the stage refilled next
      {
        const int fill = stage == 0 ? T3D_NST - 1 : stage - 1;
        if (nkk < p.nk) issue_next(fill);
        cp_async_commit();
      }
      const double* Wsm = dsm + stage * T3D_STAGE;
      const double* Tl = Wsm + 3072;
      stage = stage + 1 == T3D_NST ? 0 : stage + 1;
      const int c = TC * TT + lc;
      if (ab_ok && c < nv) {
        const double m0 = Wsm[(SWAP ? 2 : 0) * 512 + s], m1 = Wsm[(SWAP ? 4 : 1) * 512 + s];
        const double m2 = Wsm[(SWAP ? 0 : 2) * 512 + s], m3 = Wsm[(SWAP ? 5 : 3) * 512 + s];
        const double m4 = Wsm[(SWAP ? 1 : 4) * 512 + s], m5 = Wsm[(SWAP ? 3 : 5) * 512 + s];
        const double Kij_ac = Tl[0 * 64 + ac_], Kij_cb = Tl[1 * 64 + bc_], Kik_ac = Tl[2 * 64 + ac_];
        const double Kik_bc = Tl[3 * 64 + bc_], Kjk_bc = Tl[4 * 64 + bc_], Kjk_ac = Tl[5 * 64 + ac_];
        const double Tij_ac = Tl[6 * 64 + ac_], Tij_cb = Tl[7 * 64 + bc_], Tik_ac = Tl[8 * 64 + ac_];
        const double Tik_bc = Tl[9 * 64 + bc_], Tjk_bc = Tl[10 * 64 + bc_], Tjk_ac = Tl[11 * 64 + ac_];
        const double* Vs = Wsm + 3840 + lc;
        const double t1kc = Vs[0], t1jc = Vs[8], t1ic = Vs[16], fkc = Vs[24], fjc = Vs[32], fic = Vs[40];
        const double rden = 1.0 / (eabk - Vs[48]);
        const double Yn = (2.0 * Kij_ab * t1kc - Kij_ac * t1kb - Kij_cb * t1ka)
                          + (2.0 * Kik_ac * t1jb - Kik_bc * t1ja - Kik_ab * t1jc)
                          + (2.0 * Kjk_bc * t1ia - Kjk_ac * t1ib - Kjk_ba * t1ic)
                          + (2.0 * Tij_ab * fkc - Tij_ac * fkb - Tij_cb * fka)
                          + (2.0 * Tik_ac * fjb - Tik_bc * fja - Tik_ab * fjc)
                          + (2.0 * Tjk_bc * fia - Tjk_ac * fib - Tjk_ba * fic);
        const double X3 = 8.0 * m0 - 4.0 * (m1 + m2 + m5) + 2.0 * (m3 + m4);
        const double Y3 = Yn * rden;
        const double W2 = 2.0 * X3 + Y3;
        const double P = 2.0 * m0 - m1 - m5;
        const double U = m0 - m5;
        const double Z3 = 2.0 * (m0 - m1) - (m2 - m3);
        p.W2ab[oab] = W2; p.W2n[on] = W2;
        p.Pab[oab] = P;   p.Pn[on] = P;
        accG += 4.0 * t1kc * Z3;
        accX += U * fkc;
        accD += 0.5 * m0 * (X3 + Y3);
        accO += U * Tjk_bc;
        accS += (m0 - m2) * Kjk_bc;
      }
      oab += TT; on += TT;
    }
    oab += nv - (i64)nt * TT;
    on += vv - (i64)nt * TT;
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
  p.no = no; p.nv = nv; p.nt = (nv + TT - 1) / TT; p.blocked = (q_blocked & 1) ? 1 : 0; p.paired = (q_blocked & 2) ? 1 : 0;
  p.ijk = ijk; p.Q = Q; p.t1 = t1; p.t2 = t2; p.oovv = oovv; p.fov = fov; p.eo = eo; p.ev = ev; p.ldf = ldf;
  const int ncube = sorted_cubes(nv);
  // hoisted index arithmetic measured 2.47 vs 2.03 TB/s on B200 (profiles/t_probe_r01_hoist.log); B200CC_T_HOIST=0 selects
  // the per-element form
  static const bool hoist = [] { const char* e = getenv("B200CC_T_HOIST"); return !(e && e[0] == '0'); }();
  // cp.async-staged kernel (plain Q layout): B200CC_T_ENERGY=classic selects the register-load kernels below
  static const bool use_cp = [] { const char* e = getenv("B200CC_T_ENERGY"); return !(e && e[0] == 'c'); }();
  constexpr int TE_SMEM = TE_DOUBLES * (int)sizeof(double);
  static bool configured = false;
  if (!configured) {
    B200CC_CUDA_OK(cudaFuncSetAttribute(t_energy_cp_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM));
    B200CC_CUDA_OK(cudaFuncSetAttribute(t_energy_cp_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM));
    // (forcing the largest shared-memory carve-out was measured much SLOWER, 1.86 vs 2.88 TB/s: cp.async.ca wants its L1)
    configured = true;
  }
  const dim3 grid(ncube, ntrip);
  if (p.paired) {
    if (p.blocked) t_energy_kernel<true, false, 3><<<grid, 512, 0, st>>>(p, scratch);
    else if (use_cp) t_energy_cp_kernel<3><<<grid, 512, TE_SMEM, st>>>(p, scratch);
    else t_energy_kernel<false, true, 3><<<grid, 512, 0, st>>>(p, scratch);
  } else if (p.blocked) t_energy_kernel<true, false, 6><<<grid, 512, 0, st>>>(p, scratch);
  else if (use_cp) t_energy_cp_kernel<6><<<grid, 512, TE_SMEM, st>>>(p, scratch);
  else if (hoist) t_energy_kernel<false, true, 6><<<grid, 512, 0, st>>>(p, scratch);
  else t_energy_kernel<false, false, 6><<<grid, 512, 0, st>>>(p, scratch);
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
  p.no = no; p.nv = nv; p.nt = (nv + TT - 1) / TT; p.blocked = (q_blocked & 1) ? 1 : 0; p.paired = (q_blocked & 2) ? 1 : 0;
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
  p.swap_ab = d->swap_ab ? 1 : 0;
  p.M3 = d->M3; p.t1 = d->t1; p.Ts = d->t2s; p.Ks = d->oovvs; p.fov = d->fov; p.eo = d->eo; p.ev = d->ev;
  p.ldf = d->ldf; p.W2ab = d->W2ab; p.W2n = d->W2n; p.Pab = d->Pab; p.Pn = d->Pn; p.Gij = d->Gij; p.Xij = d->Xij;
  p.scratch = d->scratch;
  if ((double)d->nk * d->nv * d->nv * d->nv >= 4294967296.0 || (double)d->no * d->no * d->nv * d->nv >= 4294967296.0) {
    set_error("b200cc_t3_density_forms: nk*nv^3 and no^2*nv^2 must stay below 2^32 (shorten the k run)");
    return 1;
  }
  constexpr int SMEM = T3D_NST * T3D_STAGE * (int)sizeof(double);
  static bool configured = false;
  if (!configured) {
    B200CC_CUDA_OK(cudaFuncSetAttribute(t3_density_forms_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    B200CC_CUDA_OK(cudaFuncSetAttribute(t3_density_forms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  if (p.swap_ab) t3_density_forms_kernel<true><<<dim3(p.nt, p.nt), 512, SMEM, st>>>(p);
  else t3_density_forms_kernel<false><<<dim3(p.nt, p.nt), 512, SMEM, st>>>(p);
  if (check_launch("t3_density_forms_kernel")) return 1;
  const i64 stride = (i64)d->nv * p.nt;
  if (launch_final_reduce(d->scratch, p.nt, p.nt, d->nv, d->dvv, 1, 1.0, st)) return 1;
  if (launch_final_reduce(d->scratch + stride, p.nt, p.nt, d->nv, d->Dov, 1, 1.0, st)) return 1;
  return launch_final_reduce(d->scratch + 2 * stride, p.nt, p.nt, d->nv, d->S1, 1, 1.0, st);
}
