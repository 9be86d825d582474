// gemm.cu -- FP64 tensor-core GEMM for sm_100a (DMMA.8x8x4 via mma.sync.m8n8k4.f64).
//
// tcgen05.mma has no f64 kind, so double precision on Blackwell is the warp-level DMMA path with
// register accumulators (nvcc 12.9 lowers every f64 mma shape to DMMA.8x8x4 on sm_100a).
//
//   CTA tile 128 x 128 x 16, 8 warps as 4(M) x 2(N); each warp owns 4 x 8 interleaved 8x8 fragments
//   (fragment f of a row/column belongs to warp f % 4 / f % 2) so ragged edges are skipped at
//   8-row granularity and stay balanced across warps.  Operands are staged global->shared with a
//   4-deep cp.async pipeline (zero-fill handles every M/N/K tail); one __syncthreads per 16-wide
//   k-tile (= 128 DMMA per warp).  Either operand may be K-major or M/N-major; shared tiles are
//   padded (+4 doubles per row) so all fragment loads (LDS.64) are bank-conflict free.
//   Two K-segments can be chained into the same accumulators; split-K goes through a workspace and a
//   deterministic reduction.
#include "common.cuh"

namespace b200cc {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, NTHREADS = 256;
constexpr int WARPS_M = 4, WARPS_N = 2;
constexpr int MI = BM / (8 * WARPS_M);  // 4 M-fragments per warp
constexpr int NI = BN / (8 * WARPS_N);  // 8 N-fragments per warp
constexpr int LDK = BK + 4;             // K-major shared row pitch (doubles): 160 B -> rows 32 B apart mod 128
constexpr int LDMN = BM + 4;            // M/N-major shared row pitch: 1056 B -> rows 32 B apart mod 128
constexpr int TILE = BM * LDK;          // 2560 doubles per operand per stage (>= BK*LDMN = 2112)
constexpr int SMEM_BYTES = STAGES * 2 * TILE * (int)sizeof(double);  // 163840
static_assert(BM == BN, "shared tile helpers assume square CTA tiles");
static_assert(BK * LDMN <= TILE, "tile buffer too small");

struct KParams {
  int M, N, K1, K2;
  const double *A1, *B1, *A2, *B2;
  i64 lda1, ldb1, lda2, ldb2, sA1, sB1, sA2, sB2;
  double* C;
  i64 ldc, sC;
  double alpha, beta;
  const i64* table;
  int tiles_m, tiles_n;
  int kt1, kt_total, kt_per_split;
  int ksplit, batch, cvec;
  double* ws;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// Stage one operand tile.  TRANS=false: global (row, k) at G[row*ld + k]  -> shared [128][LDK].
//                          TRANS=true : global (k, row) at G[k*ld + row]  -> shared [16][LDMN].
// Out-of-range rows / k are zero-filled (cp.async src-size), so no tail code exists downstream.
template <bool TRANS, int VEC>
__device__ __forceinline__ void load_tile(double* sm, const double* __restrict__ G, i64 ld, int row0,
                                          int nrows, int k0, int K, int tid) {
  if (!TRANS) {
    constexpr int CPR = BK / VEC;  // chunks per row
    constexpr int NCH = BM * CPR;
#pragma unroll
    for (int i = 0; i < NCH / NTHREADS; ++i) {
      const int c = tid + i * NTHREADS;
      const int r = c / CPR, kc = (c % CPR) * VEC;
      const int gr = row0 + r, gk = k0 + kc;
      const int valid = (gr < nrows) ? min(max(K - gk, 0), VEC) : 0;
      const double* src = valid ? (G + (i64)gr * ld + gk) : G;
      const uint32_t dst = smem_u32(sm + r * LDK + kc);
      if (VEC == 2) cp_async16(dst, src, valid * 8);
      else cp_async8(dst, src, valid * 8);
    }
  } else {
    constexpr int CPR = BM / VEC;
    constexpr int NCH = BK * CPR;
#pragma unroll
    for (int i = 0; i < NCH / NTHREADS; ++i) {
      const int c = tid + i * NTHREADS;
      const int kr = c / CPR, mc = (c % CPR) * VEC;
      const int gk = k0 + kr, gm = row0 + mc;
      const int valid = (gk < K) ? min(max(nrows - gm, 0), VEC) : 0;
      const double* src = valid ? (G + (i64)gk * ld + gm) : G;
      const uint32_t dst = smem_u32(sm + kr * LDMN + mc);
      if (VEC == 2) cp_async16(dst, src, valid * 8);
      else cp_async8(dst, src, valid * 8);
    }
  }
}

// One 16-wide k-tile: 4 k4-steps x (MI x NI) DMMA per warp.
// Fragment ownership (PTX m8n8k4.f64): lane = 4*g + q holds A[g][q], B[q][g], C[g][2q], C[g][2q+1].
template <bool TA, bool TB, bool CHECK>
__device__ __forceinline__ void compute_tile(const double* __restrict__ As, const double* __restrict__ Bs,
                                             double (&acc)[MI][NI][2], int wm, int wn, int g, int q,
                                             uint32_t mmask, uint32_t nmask) {
#pragma unroll
  for (int ks = 0; ks < BK / 4; ++ks) {
    double a[MI], b[NI];
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int r = 8 * (wm + WARPS_M * i) + g;
      a[i] = TA ? As[(ks * 4 + q) * LDMN + r] : As[r * LDK + ks * 4 + q];
    }
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int r = 8 * (wn + WARPS_N * j) + g;
      b[j] = TB ? Bs[(ks * 4 + q) * LDMN + r] : Bs[r * LDK + ks * 4 + q];
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      if (CHECK && !((mmask >> i) & 1u)) continue;
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        if (CHECK && !((nmask >> j) & 1u)) continue;
        dmma(acc[i][j], a[i], b[j]);
      }
    }
  }
}

template <bool TA, bool TB, int VEC>
__global__ void __launch_bounds__(NTHREADS, 1) dgemm_kernel(const KParams p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  const int g = lane >> 2, q = lane & 3;
  const int tm = blockIdx.x % p.tiles_m, tn = blockIdx.x / p.tiles_m;
  const int m0 = tm * BM, n0 = tn * BN;
  const int b = blockIdx.y, z = blockIdx.z;

  const double *A1, *B1, *A2, *B2;
  double* C;
  if (p.table) {
    const i64* t = p.table + 5 * (i64)b;
    A1 = reinterpret_cast<const double*>(t[0]);
    B1 = reinterpret_cast<const double*>(t[1]);
    A2 = reinterpret_cast<const double*>(t[2]);
    B2 = reinterpret_cast<const double*>(t[3]);
    C = reinterpret_cast<double*>(t[4]);
  } else {
    A1 = p.A1 + (i64)b * p.sA1;
    B1 = p.B1 + (i64)b * p.sB1;
    A2 = p.A2 + (i64)b * p.sA2;
    B2 = p.B2 + (i64)b * p.sB2;
    C = p.C + (i64)b * p.sC;
  }

  const int kt_begin = z * p.kt_per_split;
  const int nkt = min(kt_begin + p.kt_per_split, p.kt_total) - kt_begin;  // may be <= 0

  uint32_t mmask = 0, nmask = 0;
#pragma unroll
  for (int i = 0; i < MI; ++i) mmask |= (m0 + 8 * (wm + WARPS_M * i) < p.M) ? (1u << i) : 0u;
#pragma unroll
  for (int j = 0; j < NI; ++j) nmask |= (n0 + 8 * (wn + WARPS_N * j) < p.N) ? (1u << j) : 0u;
  const bool full = (m0 + BM <= p.M) && (n0 + BN <= p.N);

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto issue = [&](int kt_rel, int stage) {
    const int kt = kt_begin + kt_rel;
    double* As = smem + stage * (2 * TILE);
    double* Bs = As + TILE;
    if (kt < p.kt1) {
      const int k0 = kt * BK;
      load_tile<TA, VEC>(As, A1, p.lda1, m0, p.M, k0, p.K1, tid);
      load_tile<TB, VEC>(Bs, B1, p.ldb1, n0, p.N, k0, p.K1, tid);
    } else {
      const int k0 = (kt - p.kt1) * BK;
      load_tile<TA, VEC>(As, A2, p.lda2, m0, p.M, k0, p.K2, tid);
      load_tile<TB, VEC>(Bs, B2, p.ldb2, n0, p.N, k0, p.K2, tid);
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkt) issue(s, s);
    cp_async_commit();
  }
  for (int t = 0; t < nkt; ++t) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nt = t + STAGES - 1;
    if (nt < nkt) issue(nt, nt % STAGES);
    cp_async_commit();
    const double* As = smem + (t % STAGES) * (2 * TILE);
    const double* Bs = As + TILE;
    if (full) compute_tile<TA, TB, false>(As, Bs, acc, wm, wn, g, q, mmask, nmask);
    else compute_tile<TA, TB, true>(As, Bs, acc, wm, wn, g, q, mmask, nmask);
  }
  cp_async_wait<0>();

  // ---- epilogue: C = alpha*acc + beta*C, or raw partials into the split-K workspace
  const bool split = p.ksplit > 1;
  double* out = split ? p.ws + ((i64)z * p.batch + b) * (i64)p.M * p.N : C;
  const i64 ldo = split ? (i64)p.N : p.ldc;
  const double alpha = split ? 1.0 : p.alpha, beta = split ? 0.0 : p.beta;
  const bool vec = split ? ((p.N & 1) == 0 && (((i64)p.M * p.N) & 1) == 0) : (p.cvec != 0);
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = m0 + 8 * (wm + WARPS_M * i) + g;
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int col = n0 + 8 * (wn + WARPS_N * j) + 2 * q;
      if (col >= p.N) continue;
      double* c = out + (i64)row * ldo + col;
      double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
      if (vec && col + 1 < p.N) {
        if (beta != 0.0) {
          const double2 o = *reinterpret_cast<const double2*>(c);
          v0 += beta * o.x;
          v1 += beta * o.y;
        }
        *reinterpret_cast<double2*>(c) = make_double2(v0, v1);
      } else {
        if (beta != 0.0) v0 += beta * c[0];
        c[0] = v0;
        if (col + 1 < p.N) {
          if (beta != 0.0) v1 += beta * c[1];
          c[1] = v1;
        }
      }
    }
  }
}

// C[b] = alpha * sum_z ws[z][b] + beta * C[b]
__global__ void splitk_reduce_kernel(const double* __restrict__ ws, int ksplit, int batch, int M, int N,
                                     double alpha, double beta, double* C, i64 ldc, i64 sC,
                                     const i64* table) {
  const i64 mn = (i64)M * N;
  const i64 total = mn * batch;
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    const i64 b = e / mn, r = e - b * mn;
    const i64 row = r / N, col = r - row * N;
    double s = 0.0;
    for (int z = 0; z < ksplit; ++z) s += ws[((i64)z * batch + b) * mn + r];
    double* Cb = table ? reinterpret_cast<double*>(table[5 * b + 4]) : C + b * sC;
    double* c = Cb + row * ldc + col;
    *c = (beta != 0.0) ? alpha * s + beta * (*c) : alpha * s;
  }
}

template <bool TA, bool TB, int VEC>
static int launch(const KParams& p, dim3 grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    B200CC_CUDA_OK(cudaFuncSetAttribute(dgemm_kernel<TA, TB, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        SMEM_BYTES));
    configured = true;
  }
  dgemm_kernel<TA, TB, VEC><<<grid, NTHREADS, SMEM_BYTES, st>>>(p);
  return check_launch("dgemm_kernel");
}

static inline bool even(i64 x) { return (x & 1) == 0; }
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace b200cc

using namespace b200cc;

extern "C" int b200cc_dgemm(const b200cc_gemm_desc* d, void* stream) {
  if (!d) { set_error("b200cc_dgemm: null descriptor"); return 1; }
  if (d->M < 0 || d->N < 0 || d->K1 < 0 || d->K2 < 0 || d->batch < 0) {
    set_error("b200cc_dgemm: negative dimension"); return 1;
  }
  if (d->M == 0 || d->N == 0 || d->batch == 0) return 0;
  if (d->batch > 65535) { set_error("b200cc_dgemm: batch %d > 65535 (chunk it)", d->batch); return 1; }
  const int ksplit = d->ksplit > 1 ? d->ksplit : 1;
  if (ksplit > 65535) { set_error("b200cc_dgemm: ksplit too large"); return 1; }
  if (ksplit > 1 && !d->workspace) { set_error("b200cc_dgemm: ksplit > 1 needs a workspace"); return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  KParams p;
  p.M = d->M; p.N = d->N; p.K1 = d->K1; p.K2 = d->K2;
  p.A1 = d->A1; p.B1 = d->B1;
  p.A2 = d->K2 > 0 ? d->A2 : d->A1; p.B2 = d->K2 > 0 ? d->B2 : d->B1;
  p.lda1 = d->lda1; p.ldb1 = d->ldb1;
  p.lda2 = d->K2 > 0 ? d->lda2 : d->lda1; p.ldb2 = d->K2 > 0 ? d->ldb2 : d->ldb1;
  p.sA1 = d->strideA1; p.sB1 = d->strideB1;
  p.sA2 = d->K2 > 0 ? d->strideA2 : 0; p.sB2 = d->K2 > 0 ? d->strideB2 : 0;
  p.C = d->C; p.ldc = d->ldc; p.sC = d->strideC;
  p.alpha = d->alpha; p.beta = d->beta;
  p.table = d->table;
  p.tiles_m = (d->M + BM - 1) / BM; p.tiles_n = (d->N + BN - 1) / BN;
  p.kt1 = (d->K1 + BK - 1) / BK;
  p.kt_total = p.kt1 + (d->K2 + BK - 1) / BK;
  p.ksplit = ksplit; p.batch = d->batch; p.ws = d->workspace;
  p.kt_per_split = (p.kt_total + ksplit - 1) / ksplit;
  if (p.kt_per_split < 1) p.kt_per_split = 1;

  // 16-byte vector paths need every address/pitch/stride to be a multiple of 2 doubles.
  bool va, vb, vc;
  if (d->table) {
    va = vb = vc = false;  // table entries have unknown parity unless the caller vouches for it
    if (d->table_align16) {
      va = even(p.lda1) && even(p.lda2);
      vb = even(p.ldb1) && even(p.ldb2);
      vc = even(p.ldc);
    }
  } else {
    const bool multi = d->batch > 1;
    va = al16(p.A1) && even(p.lda1) && (!multi || even(p.sA1)) &&
         (d->K2 == 0 || (al16(p.A2) && even(p.lda2) && (!multi || even(p.sA2))));
    vb = al16(p.B1) && even(p.ldb1) && (!multi || even(p.sB1)) &&
         (d->K2 == 0 || (al16(p.B2) && even(p.ldb2) && (!multi || even(p.sB2))));
    vc = al16(p.C) && even(p.ldc) && (!multi || even(p.sC));
  }
  p.cvec = vc ? 1 : 0;
  const bool v2 = va && vb;

  const i64 tiles = (i64)p.tiles_m * p.tiles_n;
  if (tiles > 2147483647LL) { set_error("b200cc_dgemm: too many tiles"); return 1; }
  dim3 grid((unsigned)tiles, (unsigned)d->batch, (unsigned)ksplit);
  const int ta = d->transA ? 1 : 0, tb = d->transB ? 1 : 0;
  int rc;
#define B200CC_GO(TA, TB)                                             \
  rc = v2 ? launch<TA, TB, 2>(p, grid, st) : launch<TA, TB, 1>(p, grid, st)
  if (!ta && !tb) { B200CC_GO(false, false); }
  else if (!ta && tb) { B200CC_GO(false, true); }
  else if (ta && !tb) { B200CC_GO(true, false); }
  else { B200CC_GO(true, true); }
#undef B200CC_GO
  if (rc) return rc;
  if (ksplit > 1) {
    const i64 total = (i64)d->M * d->N * d->batch;
    int blocks = (int)((total + 255) / 256);
    const int cap = sm_count() * 8;
    if (blocks > cap) blocks = cap;
    splitk_reduce_kernel<<<blocks, 256, 0, st>>>(d->workspace, ksplit, d->batch, d->M, d->N, d->alpha, d->beta,
                                                 d->C, d->ldc, d->strideC, d->table);
    return check_launch("splitk_reduce_kernel");
  }
  return 0;
}
