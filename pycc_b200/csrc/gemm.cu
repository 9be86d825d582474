// gemm.cu -- FP64 tensor-core GEMM for sm_100a (DMMA.8x8x4 via mma.sync.m8n8k4.f64).
//
// tcgen05.mma has no f64 kind, so double precision on Blackwell is the warp-level DMMA path with
// register accumulators (nvcc 12.9 lowers every f64 mma shape to DMMA.8x8x4 on sm_100a; one DMMA
// occupies an SM sub-partition's tensor pipe for 16 cycles = 64 FMA/clk/SM).
//
//   * PERSISTENT: one CTA per SM walks a list of work units (output tile x batch entry x K-split);
//     the 4-deep cp.async operand pipeline runs ACROSS unit boundaries, so the next tile's operands
//     stream in while the current tile is finished and stored (no per-tile prologue/epilogue bubble --
//     this matters for the short-K GEMMs of (T) and of the o^3v^3 ring terms).
//   * CTA tile BM x BN x 16 from a compile-time config <WARPS_M, WARPS_N, MI, NI>; each warp owns
//     MI x NI INTERLEAVED 8x8 fragments (fragment f of a row/column belongs to warp f % WARPS), so ragged
//     edges are skipped at 8-row granularity and stay balanced over warps and SM sub-partitions.
//   * Either operand K-major or M/N-major; shared tiles are padded (+4 doubles per row) so every
//     fragment load (LDS.64) is bank-conflict free; cp.async zero-fill handles all M/N/K tails.
//   * Two K-segments can be chained into the same accumulators ((T): particle + hole term);
//     split-K goes through a workspace and a deterministic reduction; batches by stride or address table.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "pipe.cuh"

namespace b200cc {

constexpr int BK = 16, STAGES = 4;
constexpr int FULL_CODE = 99;   // consume_unit: every fragment of the warp is in range
constexpr int LDK = BK + 4;  // K-major shared row pitch (doubles): 160 B -> consecutive rows 32 B apart mod 128

template <int WM_, int WN_, int MI_, int NI_>
struct Cfg {
  static constexpr int WARPS_M = WM_, WARPS_N = WN_, MI = MI_, NI = NI_;
  static constexpr int BM = 8 * WM_ * MI_, BN = 8 * WN_ * NI_, NT = 32 * WM_ * WN_;
  static constexpr int TILE_A = (BM * LDK > BK * (BM + 4)) ? BM * LDK : BK * (BM + 4);
  static constexpr int TILE_B = (BN * LDK > BK * (BN + 4)) ? BN * LDK : BK * (BN + 4);
  static constexpr int STAGE = TILE_A + TILE_B;
  static constexpr int SMEM_BYTES = STAGES * STAGE * (int)sizeof(double);
};
typedef Cfg<4, 2, 4, 8> CfgB;  // 128 x 128,  8 warps
typedef Cfg<2, 4, 5, 4> CfgC;  //  80 x 128,  8 warps (M = o^2 = 400, 1600, ... divide by 80)
// tall-skinny products (one extent = o <= 40: the t1 contractions with <mb|ef>, Fae / Fmi builds, r1 terms)
typedef Cfg<1, 8, 5, 4> CfgD;  //  40 x 256
typedef Cfg<8, 1, 4, 5> CfgE;  // 256 x  40

struct KParams {
  int M, N, K1, K2;
  const double *A1, *B1, *A2, *B2;
  i64 lda1, ldb1, lda2, ldb2, sA1, sB1, sA2, sB2;
  double* C;
  i64 ldc, sC;
  double alpha, beta;
  const i64* table;
  int tiles_m, tiles_n, tiles;
  int kt1, kt_total, kt_per_split;
  int ksplit, batch, cvec, nfast;
  int units;
  double* ws;
  const int* bcoords;
  int nbA1, nbB1, nbA2, nbB2;
  int cube_nv;   // > 0: C written as contiguous 8x8x8 cubes of (x = row / nv, y = row % nv, z = col)
  // third / fourth K segment (TMA kernels only): kt2, kt3 = cumulative k-tile counts after segments 2 and 3
  int K3, K4, kt2, kt3, bc_stride;
  int tile0, compact;   // split-K of the LAST wave only: units cover output tiles [tile0, tiles*batch), partials in compact tile slots
  const double *A3, *B3, *A4, *B4;
  i64 lda3, ldb3, lda4, ldb4, sA3, sB3, sA4, sB4;
  int nbA3, nbB3, nbA4, nbB4;
};

struct alignas(64) TMapSet {
  CUtensorMap a[4], b[4];
};

// Stage one operand tile of ROWS rows.  TRANS=false: global (row,k) at G[row*ld + k] -> shared [ROWS][LDK].
//                                       TRANS=true : global (k,row) at G[k*ld + row] -> shared [BK][ROWS+4].
// Out-of-range rows / k are zero-filled (cp.async src-size), so no tail code exists downstream.
template <bool TRANS, int VEC, int ROWS, int NT>
__device__ __forceinline__ void load_tile(double* sm, const double* __restrict__ G, i64 ld, int row0,
                                          int nrows, int k0, int K, int tid) {
  if (!TRANS) {
    constexpr int CPR = BK / VEC;
    constexpr int NCH = ROWS * CPR;
#pragma unroll
    for (int i = 0; i < (NCH + NT - 1) / NT; ++i) {
      const int c = tid + i * NT;
      if (NCH % NT != 0 && c >= NCH) break;
      const int r = c / CPR, kc = (c % CPR) * VEC;
      const int gr = row0 + r, gk = k0 + kc;
      const int valid = (gr < nrows) ? min(max(K - gk, 0), VEC) : 0;
      const double* src = valid ? (G + (i64)gr * ld + gk) : G;
      const uint32_t dst = smem_u32(sm + r * LDK + kc);
      if (VEC == 2) cp_async16(dst, src, valid * 8);
      else cp_async8(dst, src, valid * 8);
    }
  } else {
    constexpr int LDMN = ROWS + 4;
    constexpr int CPR = ROWS / VEC;
    constexpr int NCH = BK * CPR;
#pragma unroll
    for (int i = 0; i < (NCH + NT - 1) / NT; ++i) {
      const int c = tid + i * NT;
      if (NCH % NT != 0 && c >= NCH) break;
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
template <class CF, bool TA, bool TB, bool CHECK>
__device__ __forceinline__ void compute_tile(const double* __restrict__ As, const double* __restrict__ Bs,
                                             double (&acc)[CF::MI][CF::NI][2], int wm, int wn, int g, int q,
                                             uint32_t mmask, uint32_t nmask) {
  constexpr int MI = CF::MI, NI = CF::NI;
#pragma unroll
  for (int ks = 0; ks < BK / 4; ++ks) {
    double a[MI], b[NI];
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int r = 8 * (wm + CF::WARPS_M * i) + g;
      a[i] = TA ? As[(ks * 4 + q) * (CF::BM + 4) + r] : As[r * LDK + ks * 4 + q];
    }
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int r = 8 * (wn + CF::WARPS_N * j) + g;
      b[j] = TB ? Bs[(ks * 4 + q) * (CF::BN + 4) + r] : Bs[r * LDK + ks * 4 + q];
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

// Persistent scheduling (one CTA per SM striding over the units) pipelines across tile boundaries, which pays
// for short-K units.  For long-K units it buys nothing, and the fixed unit->CTA map lets co-running CTAs drift
// apart in k, so operand panels shared by neighbouring tiles stop meeting in L2 (ncu: 1.09 TB of DRAM reads for the
// o=40,v=300 ladder, 16x the algorithmic bytes).  Long-K GEMMs therefore launch one CTA per unit: the hardware
// hands out units in order as SMs free up, so neighbouring units start together.
static inline int pick_grid(const KParams& p) {
  const int nsm = sm_count();
  if (p.units <= nsm) return p.units;
  const int force = getenv("B200CC_GEMM_PERSISTENT") ? atoi(getenv("B200CC_GEMM_PERSISTENT")) : -1;
  if (force == 1) return nsm;
  if (force == 0) return p.units;
  return p.kt_per_split >= 512 ? p.units : nsm;
}

struct Unit {
  int m0, n0, b, z, kt_begin, nkt, tidx;
};
struct LState {
  const double *A1, *B1, *A2, *B2;
  int m0, n0, kt_begin, pad;
};

template <class CF>
__device__ __forceinline__ Unit decode_unit(const KParams& p, int u) {
  Unit w;
  const int per = p.tiles * p.batch - p.tile0;      // output tiles covered by this launch (all of them unless tile0 > 0)
  w.z = u / per;
  w.tidx = u - w.z * per;
  const int rem = p.tile0 + w.tidx;
  w.b = rem / p.tiles;
  const int t = rem - w.b * p.tiles;
  int tm, tn;
  if (p.nfast) { tn = t % p.tiles_n; tm = t / p.tiles_n; }
  else { tm = t % p.tiles_m; tn = t / p.tiles_m; }
  w.m0 = tm * CF::BM;
  w.n0 = tn * CF::BN;
  w.kt_begin = w.z * p.kt_per_split;
  const int e = min(w.kt_begin + p.kt_per_split, p.kt_total);
  w.nkt = e > w.kt_begin ? e - w.kt_begin : 0;
  return w;
}

// Epilogue of one work unit: C = alpha*acc + beta*C (or raw partials into the split-K workspace).  With beta != 0 the old
// values of a whole fragment row are loaded BEFORE the first store of that row: load -> fma -> store pairs in program order
// serialise on the memory latency (the store waits for its load, the next load issues behind the store) -- measured
// 20 us per 128 x 128 tile, 1.5 ms of a 4.0 ms short-K product with 11 250 tiles.
template <class CF, bool TABLE>
__device__ __forceinline__ void gemm_epilogue(const double (&acc)[CF::MI][CF::NI][2], const KParams& p, const Unit& w,
                                              int wm, int wn, int g, int q) {
  constexpr int MI = CF::MI, NI = CF::NI;
  const bool split = p.ksplit > 1;
  const bool compact = split && p.compact != 0;     // partial tile (z, tidx) as a dense BM x BN block of the workspace
  double* C;
  if (compact) C = p.ws + ((i64)w.z * (p.tiles * p.batch - p.tile0) + w.tidx) * (i64)(CF::BM * CF::BN) -
                   ((i64)w.m0 * CF::BN + w.n0);
  else if (split) C = p.ws + ((i64)w.z * p.batch + w.b) * (i64)p.M * p.N;
  else if (TABLE && p.table) C = reinterpret_cast<double*>(p.table[5 * (i64)w.b + 4]);
  else C = p.C + (i64)w.b * p.sC;
  const i64 ldo = compact ? (i64)CF::BN : (split ? (i64)p.N : p.ldc);
  const double alpha = split ? 1.0 : p.alpha, beta = split ? 0.0 : p.beta;
  const bool cube = !split && p.cube_nv > 0;
  const bool vec = compact ? true : (split ? ((p.N & 1) == 0 && (((i64)p.M * p.N) & 1) == 0) : (p.cvec != 0 || cube));
  const int nc8 = (p.cube_nv + 7) >> 3;
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = w.m0 + 8 * (wm + CF::WARPS_M * i) + g;
    if (row >= p.M) continue;
    const int cx = cube ? row / p.cube_nv : 0, cy = cube ? row - cx * p.cube_nv : 0;
    // JC fragment columns at a time: JC loads in flight, then JC stores (a whole row of 8 would cost 48 more registers)
    constexpr int JC = NI >= 8 ? 4 : NI;
#pragma unroll
    for (int j0 = 0; j0 < NI; j0 += JC) {
      double* cp[JC];
      double o0[JC], o1[JC];
#pragma unroll
      for (int jj = 0; jj < JC; ++jj) {
        const int j = j0 + jj;
        const int col = w.n0 + 8 * (wn + CF::WARPS_N * j) + 2 * q;
        // (T): Q stored as 4 KB cubes so that the energy kernel reads it in fully contiguous runs
        cp[jj] = cube ? C + ((((i64)(cx >> 3) * nc8 + (cy >> 3)) * nc8 + (col >> 3)) << 9) + ((cx & 7) << 6) +
                            ((cy & 7) << 3) + (col & 7)
                      : C + (i64)row * ldo + col;
        o0[jj] = o1[jj] = 0.0;
        if (beta != 0.0 && col < p.N) {
          if (vec && col + 1 < p.N) {
            const double2 o = *reinterpret_cast<const double2*>(cp[jj]);
            o0[jj] = o.x;
            o1[jj] = o.y;
          } else {
            o0[jj] = cp[jj][0];
            if (col + 1 < p.N) o1[jj] = cp[jj][1];
          }
        }
      }
#pragma unroll
      for (int jj = 0; jj < JC; ++jj) {
        const int j = j0 + jj;
        const int col = w.n0 + 8 * (wn + CF::WARPS_N * j) + 2 * q;
        if (j >= NI || col >= p.N) continue;
        double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
        if (beta != 0.0) {
          v0 += beta * o0[jj];
          v1 += beta * o1[jj];
        }
        if (vec && col + 1 < p.N) {
          *reinterpret_cast<double2*>(cp[jj]) = make_double2(v0, v1);
        } else {
          cp[jj][0] = v0;
          if (col + 1 < p.N) cp[jj][1] = v1;
        }
      }
    }
  }
}

template <class CF, bool TA, bool TB, int VEC>
__global__ void __launch_bounds__(CF::NT, 1) dgemm_kernel(const KParams p) {
  extern __shared__ __align__(16) double smem[];
  constexpr int MI = CF::MI, NI = CF::NI;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wn = warp % CF::WARPS_N, wm = warp / CF::WARPS_N;
  const int g = lane >> 2, q = lane & 3;
  const int G = gridDim.x;

  // ---- load cursor: (unit, k-tile) of the next operand tile to stage; skips units with no k-tiles.
  // The decoded unit (tile origin, operand base pointers, k range) lives in shared memory, double
  // buffered, and is rewritten only when the cursor enters a new unit -- the per-k-tile path has no
  // integer divisions.  Every issue_next() is preceded by a __syncthreads(), which orders the
  // (thread 0) write of slot s^1 before all reads of it and after all reads of its previous content.
  __shared__ LState ls[2];
  int lu = blockIdx.x, lkt = 0, lnkt = 0, lslot = 1;
  auto seek = [&]() {
    Unit w;
    w.nkt = 0;
    while (lu < p.units) {
      w = decode_unit<CF>(p, lu);
      if (w.nkt > 0) break;
      lu += G;
    }
    lkt = 0;
    lnkt = w.nkt;
    lslot ^= 1;
    if (tid == 0 && lu < p.units) {
      LState& d = ls[lslot];
      if (p.table) {
        const i64* t = p.table + 5 * (i64)w.b;
        d.A1 = reinterpret_cast<const double*>(t[0]);
        d.B1 = reinterpret_cast<const double*>(t[1]);
        d.A2 = reinterpret_cast<const double*>(t[2]);
        d.B2 = reinterpret_cast<const double*>(t[3]);
      } else {
        d.A1 = p.A1 + (i64)w.b * p.sA1;
        d.B1 = p.B1 + (i64)w.b * p.sB1;
        d.A2 = p.A2 + (i64)w.b * p.sA2;
        d.B2 = p.B2 + (i64)w.b * p.sB2;
      }
      d.m0 = w.m0;
      d.n0 = w.n0;
      d.kt_begin = w.kt_begin;
    }
  };
  auto issue_next = [&](int stage) {
    if (lu < p.units) {
      const LState& d = ls[lslot];
      double* As = smem + stage * CF::STAGE;
      double* Bs = As + CF::TILE_A;
      const int kt = d.kt_begin + lkt;
      if (kt < p.kt1) {
        const int k0 = kt * BK;
        load_tile<TA, VEC, CF::BM, CF::NT>(As, d.A1, p.lda1, d.m0, p.M, k0, p.K1, tid);
        load_tile<TB, VEC, CF::BN, CF::NT>(Bs, d.B1, p.ldb1, d.n0, p.N, k0, p.K1, tid);
      } else {
        const int k0 = (kt - p.kt1) * BK;
        load_tile<TA, VEC, CF::BM, CF::NT>(As, d.A2, p.lda2, d.m0, p.M, k0, p.K2, tid);
        load_tile<TB, VEC, CF::BN, CF::NT>(Bs, d.B2, p.ldb2, d.n0, p.N, k0, p.K2, tid);
      }
      if (++lkt == lnkt) {
        lu += G;
        seek();
      }
    }
    cp_async_commit();
  };

  seek();
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    __syncthreads();
    issue_next(s);
  }
  int cstage = 0, lstage = STAGES - 1;

  for (int cu = blockIdx.x; cu < p.units; cu += G) {
    const Unit w = decode_unit<CF>(p, cu);
    uint32_t mmask = 0, nmask = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i) mmask |= (w.m0 + 8 * (wm + CF::WARPS_M * i) < p.M) ? (1u << i) : 0u;
#pragma unroll
    for (int j = 0; j < NI; ++j) nmask |= (w.n0 + 8 * (wn + CF::WARPS_N * j) < p.N) ? (1u << j) : 0u;
    const bool full = (w.m0 + CF::BM <= p.M) && (w.n0 + CF::BN <= p.N);

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int t = 0; t < w.nkt; ++t) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      issue_next(lstage);
      lstage = (lstage + 1 == STAGES) ? 0 : lstage + 1;
      const double* As = smem + cstage * CF::STAGE;
      const double* Bs = As + CF::TILE_A;
      if (full) compute_tile<CF, TA, TB, false>(As, Bs, acc, wm, wn, g, q, mmask, nmask);
      else compute_tile<CF, TA, TB, true>(As, Bs, acc, wm, wn, g, q, mmask, nmask);
      cstage = (cstage + 1 == STAGES) ? 0 : cstage + 1;
    }

    // ---- epilogue: C = alpha*acc + beta*C, or raw partials into the split-K workspace
    gemm_epilogue<CF, true>(acc, p, w, wm, wn, g, q);
  }
  cp_async_wait<0>();
}

// =====================================================================================================
// Warp-specialised variant (config 4): 8 consumer warps (32x64 warp tiles, the CfgB geometry) that execute
// nothing but LDS.64 + DMMA, and 4 producer warps that stage the operand tiles with cp.async and signal
// per-stage mbarriers (cp.async.mbarrier.arrive).  No __syncthreads in the main loop: a consumer warp only
// ever waits for "stage s is full", a producer only for "stage s is empty", so the warps of one SM
// sub-partition drift apart and one warp's bookkeeping hides behind the other's DMMAs (each DMMA holds
// its warp for 16 cycles anyway).  Consumers keep fragments double-buffered across k-steps and across the
// full-barrier wait of the next k-tile, so the tensor pipe does not drain at tile boundaries.
// Registers are re-partitioned with setmaxnreg (producers 64, consumers 216: 128*64 + 256*216 <= 384*168, the CTA's launch allocation -- the pool setmaxnreg draws from).
// =====================================================================================================
constexpr int WS_PRODUCER_WARPS = 4;
constexpr int WS_PRODUCER_THREADS = 32 * WS_PRODUCER_WARPS;

// Ragged tiles: a predicated-off DMMA still occupies the tensor pipe for its 16 cycles (measured with ncu:
// pipe-active time of the N = 280 (T) GEMM equalled 3 FULL N tiles), so out-of-range fragments must be
// skipped by real, warp-uniform control flow.  In-range fragments of a warp form a prefix (fragments are
// interleaved over the warps); `code` selects one of 16 straight-line DMMA blocks:
//   rows  : all MI fragments, or the first ceil(MI/2)   (code >> 3)
//   cols  : the first 1 .. NI fragments                 ((code & 7) + 1)
template <class CF, bool TA, bool TB>
__device__ __forceinline__ void load_frags(const double* __restrict__ As, const double* __restrict__ Bs, int ks,
                                           double (&a)[CF::MI], double (&b)[CF::NI], int wm, int wn, int g, int q) {
#pragma unroll
  for (int i = 0; i < CF::MI; ++i) {
    const int r = 8 * (wm + CF::WARPS_M * i) + g;
    a[i] = TA ? As[(ks * 4 + q) * (CF::BM + 4) + r] : As[r * LDK + ks * 4 + q];
  }
#pragma unroll
  for (int j = 0; j < CF::NI; ++j) {
    const int r = 8 * (wn + CF::WARPS_N * j) + g;
    b[j] = TB ? Bs[(ks * 4 + q) * (CF::BN + 4) + r] : Bs[r * LDK + ks * 4 + q];
  }
}

template <class CF, int MC, int NC>
__device__ __forceinline__ void mma_block(double (&acc)[CF::MI][CF::NI][2], const double (&a)[CF::MI],
                                          const double (&b)[CF::NI]) {
#pragma unroll
  for (int i = 0; i < MC; ++i)
#pragma unroll
    for (int j = 0; j < NC; ++j) dmma(acc[i][j], a[i], b[j]);
}

template <class CF>
__device__ __forceinline__ void mma_select(double (&acc)[CF::MI][CF::NI][2], const double (&a)[CF::MI],
                                           const double (&b)[CF::NI], int code) {
  constexpr int MI = CF::MI, MH = (CF::MI + 1) / 2, NI = CF::NI;
  static_assert(NI == 4 || NI == 5 || NI == 8, "mma_select is written for NI = 4, 5 or 8");
  // code = 8 * (rows: 0 = all MI, 1 = first ceil(MI/2)) + (in-range N fragments - 1)
#define B200CC_ARM(C, MC, NC) \
  case C: mma_block<CF, MC, (NC <= NI ? NC : NI)>(acc, a, b); break;
  switch (code) {
    B200CC_ARM(0, MI, 1) B200CC_ARM(1, MI, 2) B200CC_ARM(2, MI, 3) B200CC_ARM(3, MI, 4)
    B200CC_ARM(4, MI, 5) B200CC_ARM(5, MI, 6) B200CC_ARM(6, MI, 7) B200CC_ARM(7, MI, 8)
    B200CC_ARM(8, MH, 1) B200CC_ARM(9, MH, 2) B200CC_ARM(10, MH, 3) B200CC_ARM(11, MH, 4)
    B200CC_ARM(12, MH, 5) B200CC_ARM(13, MH, 6) B200CC_ARM(14, MH, 7)
    default: mma_block<CF, MH, NI>(acc, a, b); break;
  }
#undef B200CC_ARM
}

// The k-loop of one work unit for one consumer warp.  Fragments are double-buffered across the four
// k-steps of a tile and across the full-barrier wait of the next tile.
template <class CF, bool TA, bool TB>
__device__ __forceinline__ void consume_unit(double (&acc)[CF::MI][CF::NI][2], const double* smem,
                                             uint64_t* full_bar, uint64_t* empty_bar, int& stage, uint32_t& phase,
                                             int nkt, int code, int wm, int wn, int g, int q, int lane) {
  double a0[CF::MI], b0[CF::NI], a1[CF::MI], b1[CF::NI];
  if (nkt > 0) {
    mbar_wait(full_bar + stage, phase);
    load_frags<CF, TA, TB>(smem + stage * CF::STAGE, smem + stage * CF::STAGE + CF::TILE_A, 0, a0, b0, wm, wn, g, q);
  }
#define B200CC_KLOOP(MMA)                                                                              \
  for (int t = 0; t < nkt; ++t) {                                                                      \
    const double* As = smem + stage * CF::STAGE;                                                       \
    const double* Bs = As + CF::TILE_A;                                                                \
    load_frags<CF, TA, TB>(As, Bs, 1, a1, b1, wm, wn, g, q);                                           \
    MMA(a0, b0);                                                                                       \
    load_frags<CF, TA, TB>(As, Bs, 2, a0, b0, wm, wn, g, q);                                           \
    MMA(a1, b1);                                                                                       \
    load_frags<CF, TA, TB>(As, Bs, 3, a1, b1, wm, wn, g, q);                                           \
    MMA(a0, b0);                                                                                       \
    int nstage = stage + 1;                                                                            \
    uint32_t nphase = phase;                                                                           \
    if (nstage == STAGES) { nstage = 0; nphase ^= 1u; }                                                \
    if (t + 1 < nkt) {                                                                                 \
      mbar_wait(full_bar + nstage, nphase);                                                            \
      load_frags<CF, TA, TB>(smem + nstage * CF::STAGE, smem + nstage * CF::STAGE + CF::TILE_A, 0, a0, b0, wm, \
                             wn, g, q);                                                                \
    }                                                                                                  \
    MMA(a1, b1);                                                                                       \
    /* every shared read of `stage` has been consumed by a DMMA above: hand the slot back */           \
    __syncwarp();                                                                                      \
    if (lane == 0) mbar_arrive(empty_bar + stage);                                                     \
    stage = nstage;                                                                                    \
    phase = nphase;                                                                                    \
  }
#define B200CC_MMA_FULL(A_, B_) mma_block<CF, CF::MI, CF::NI>(acc, A_, B_)
#define B200CC_MMA_SEL(A_, B_) mma_select<CF>(acc, A_, B_, code)
  if (code == FULL_CODE) {
    B200CC_KLOOP(B200CC_MMA_FULL)      // full tile: branch-free hot loop
  } else {
    B200CC_KLOOP(B200CC_MMA_SEL)
  }
#undef B200CC_KLOOP
#undef B200CC_MMA_FULL
#undef B200CC_MMA_SEL
}

// a warp with no in-range fragment in this unit still has to keep the barriers moving
__device__ __forceinline__ void idle_unit(uint64_t* full_bar, uint64_t* empty_bar, int& stage, uint32_t& phase,
                                          int nkt, int lane) {
  for (int t = 0; t < nkt; ++t) {
    mbar_wait(full_bar + stage, phase);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar + stage);
    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
  }
}

template <class CF, bool TA, bool TB, int VEC>
__global__ void __launch_bounds__(CF::NT + WS_PRODUCER_THREADS, 1) dgemm_ws_kernel(const KParams p) {
  extern __shared__ __align__(16) double smem[];
  constexpr int MI = CF::MI, NI = CF::NI;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * CF::STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, WS_PRODUCER_THREADS);
      mbar_init(empty_bar + s, CF::NT / 32);
    }
  }
  __syncthreads();

  if (warp >= CF::NT / 32) {
    // ================================ producers ================================
    reg_dec<64>();
    const int ptid = tid - CF::NT;
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < p.units; u += G) {
      const Unit w = decode_unit<CF>(p, u);
      if (w.nkt == 0) continue;
      const double *A1, *B1, *A2, *B2;
      if (p.table) {
        const i64* t = p.table + 5 * (i64)w.b;
        A1 = reinterpret_cast<const double*>(t[0]);
        B1 = reinterpret_cast<const double*>(t[1]);
        A2 = reinterpret_cast<const double*>(t[2]);
        B2 = reinterpret_cast<const double*>(t[3]);
      } else {
        A1 = p.A1 + (i64)w.b * p.sA1;
        B1 = p.B1 + (i64)w.b * p.sB1;
        A2 = p.A2 + (i64)w.b * p.sA2;
        B2 = p.B2 + (i64)w.b * p.sB2;
      }
      for (int t = 0; t < w.nkt; ++t) {
        mbar_wait(empty_bar + stage, phase ^ 1u);      // slot free (passes immediately on the first lap)
        double* As = smem + stage * CF::STAGE;
        double* Bs = As + CF::TILE_A;
        const int kt = w.kt_begin + t;
        if (kt < p.kt1) {
          const int k0 = kt * BK;
          load_tile<TA, VEC, CF::BM, WS_PRODUCER_THREADS>(As, A1, p.lda1, w.m0, p.M, k0, p.K1, ptid);
          load_tile<TB, VEC, CF::BN, WS_PRODUCER_THREADS>(Bs, B1, p.ldb1, w.n0, p.N, k0, p.K1, ptid);
        } else {
          const int k0 = (kt - p.kt1) * BK;
          load_tile<TA, VEC, CF::BM, WS_PRODUCER_THREADS>(As, A2, p.lda2, w.m0, p.M, k0, p.K2, ptid);
          load_tile<TB, VEC, CF::BN, WS_PRODUCER_THREADS>(Bs, B2, p.ldb2, w.n0, p.N, k0, p.K2, ptid);
        }
        mbar_arrive_cp_async(full_bar + stage);         // fires when this thread's copies have landed
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    cp_async_wait<0>();
    return;
  }

  // ================================ consumers ================================
  reg_inc<216>();
  const int wn = warp % CF::WARPS_N, wm = warp / CF::WARPS_N;
  const int g = lane >> 2, q = lane & 3;
  int stage = 0;
  uint32_t phase = 0;
  for (int cu = blockIdx.x; cu < p.units; cu += G) {
    const Unit w = decode_unit<CF>(p, cu);
    // in-range fragments of this warp form a prefix (fragments are interleaved over the warps)
    int mc = 0, nc = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i) mc += (w.m0 + 8 * (wm + CF::WARPS_M * i) < p.M) ? 1 : 0;
#pragma unroll
    for (int j = 0; j < NI; ++j) nc += (w.n0 + 8 * (wn + CF::WARPS_N * j) < p.N) ? 1 : 0;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (mc == 0 || nc == 0) {
      idle_unit(full_bar, empty_bar, stage, phase, w.nkt, lane);
    } else {
      constexpr int MH = (MI + 1) / 2;
      const int code = (mc == MI && nc == NI) ? FULL_CODE : (mc <= MH ? 8 : 0) + nc - 1;
      consume_unit<CF, TA, TB>(acc, smem, full_bar, empty_bar, stage, phase, w.nkt, code, wm, wn, g, q, lane);
    }

    // ---- epilogue: C = alpha*acc + beta*C, or raw partials into the split-K workspace
    gemm_epilogue<CF, true>(acc, p, w, wm, wn, g, q);
  }
}

template <class CF, bool TA, bool TB, int VEC>
static int launch_ws(const KParams& p, int grid, cudaStream_t st) {
  constexpr int SMEM = CF::SMEM_BYTES + 2 * STAGES * (int)sizeof(uint64_t);
  static bool configured = false;
  if (!configured) {
    B200CC_CUDA_OK(cudaFuncSetAttribute(dgemm_ws_kernel<CF, TA, TB, VEC>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  dgemm_ws_kernel<CF, TA, TB, VEC><<<grid, CF::NT + WS_PRODUCER_THREADS, SMEM, st>>>(p);
  return check_launch("dgemm_ws_kernel");
}

// =====================================================================================================
// TMA variant (config 6) of the warp-specialised kernel, for K-major x K-major operands (the ladder, the
// W_mbej builds, Z_mbij, W_mnij): ONE producer thread issues two cp.async.bulk.tensor loads per stage
// (128 rows x 16 doubles = 128-byte rows, CU_TENSOR_MAP_SWIZZLE_128B, out-of-range rows/k zero-filled by
// the TMA unit) that complete on the stage's mbarrier (expect_tx); the 8 consumer warps are unchanged
// except for the fragment addressing.  With 128-byte rows a naive fragment read would be an 8-way bank
// conflict; the 128B swizzle XORs the 16-byte chunk index with (row & 7), and because the summation index
// may be visited in any order, k-step s / lane q reads logical k = 8*(q>>1) + 2*s + (q&1): the half-warp
// then covers all eight chunks x both halves -> conflict-free LDS.64 (same mapping for A and B).
// =====================================================================================================
template <class CF>
__device__ __forceinline__ void load_frags_swz(const unsigned char* __restrict__ As, const unsigned char* __restrict__ Bs,
                                               int off, double (&a)[CF::MI], double (&b)[CF::NI], int wm, int wn,
                                               int g) {
#pragma unroll
  for (int i = 0; i < CF::MI; ++i)
    a[i] = *reinterpret_cast<const double*>(As + (8 * (wm + CF::WARPS_M * i) + g) * 128 + off);
#pragma unroll
  for (int j = 0; j < CF::NI; ++j)
    b[j] = *reinterpret_cast<const double*>(Bs + (8 * (wn + CF::WARPS_N * j) + g) * 128 + off);
}

template <class CF>
__global__ void __launch_bounds__(CF::NT + WS_PRODUCER_THREADS, 1)
    dgemm_tma_kernel(const KParams p, const __grid_constant__ TMapSet tm) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int MI = CF::MI, NI = CF::NI;
  constexpr int A_BYTES = CF::BM * 128, B_BYTES = CF::BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte swizzle atoms
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, CF::NT / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp >= CF::NT / 32) {
    // ================================ producer: one thread drives the TMA unit ================================
    // (a whole warpgroup is launched only so that setmaxnreg can hand its registers to the consumers)
    reg_dec<40>();
    if (warp == CF::NT / 32 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < p.units; u += G) {
        const Unit w = decode_unit<CF>(p, u);
        int bA[4] = {p.sA1 ? w.b : 0, p.sA2 ? w.b : 0, p.sA3 ? w.b : 0, p.sA4 ? w.b : 0};
        int bB[4] = {p.sB1 ? w.b : 0, p.sB2 ? w.b : 0, p.sB3 ? w.b : 0, p.sB4 ? w.b : 0};
        if (p.bcoords) {   // (T): each batch entry names its own slab of every operand: {A1,B1,A2,B2[,A3,B3,A4,B4]}
          const int4 c = reinterpret_cast<const int4*>(p.bcoords + (i64)w.b * p.bc_stride)[0];
          bA[0] = c.x; bB[0] = c.y; bA[1] = c.z; bB[1] = c.w;
          if (p.bc_stride == 8) {
            const int4 e = reinterpret_cast<const int4*>(p.bcoords + (i64)w.b * 8)[1];
            bA[2] = e.x; bB[2] = e.y; bA[3] = e.z; bB[3] = e.w;
          }
        }
        for (int t = 0; t < w.nkt; ++t) {
          mbar_wait(empty_bar + stage, phase ^ 1u);
          unsigned char* As = tiles + stage * STAGE_BYTES;
          unsigned char* Bs = As + A_BYTES;
          mbar_expect_tx(full_bar + stage, STAGE_BYTES);
          const int kt = w.kt_begin + t;
          const int seg = (kt >= p.kt1 ? 1 : 0) + (kt >= p.kt2 ? 1 : 0) + (kt >= p.kt3 ? 1 : 0);
          const int k0 = (kt - (seg == 0 ? 0 : (seg == 1 ? p.kt1 : (seg == 2 ? p.kt2 : p.kt3)))) * BK;
          tma_load_3d(As, &tm.a[seg], k0, w.m0, bA[seg], full_bar + stage);
          tma_load_3d(Bs, &tm.b[seg], k0, w.n0, bB[seg], full_bar + stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;
  }

  // ================================ consumers ================================
  reg_inc<216>();
  const int wn = warp % CF::WARPS_N, wm = warp / CF::WARPS_N;
  const int g = lane >> 2, q = lane & 3;
  int off[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) off[s] = ((((q >> 1) * 4 + s) ^ g) << 4) + ((q & 1) << 3);
  int stage = 0;
  uint32_t phase = 0;
  for (int cu = blockIdx.x; cu < p.units; cu += G) {
    const Unit w = decode_unit<CF>(p, cu);
    int mc = 0, nc = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i) mc += (w.m0 + 8 * (wm + CF::WARPS_M * i) < p.M) ? 1 : 0;
#pragma unroll
    for (int j = 0; j < NI; ++j) nc += (w.n0 + 8 * (wn + CF::WARPS_N * j) < p.N) ? 1 : 0;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (mc == 0 || nc == 0) {
      idle_unit(full_bar, empty_bar, stage, phase, w.nkt, lane);
    } else {
      constexpr int MH = (MI + 1) / 2;
      const int code = (mc <= MH ? 8 : 0) + nc - 1;
      double a0[MI], b0[NI], a1[MI], b1[NI];
      if (w.nkt > 0) {
        mbar_wait(full_bar + stage, phase);
        load_frags_swz<CF>(tiles + stage * STAGE_BYTES, tiles + stage * STAGE_BYTES + A_BYTES, off[0], a0, b0, wm, wn, g);
      }
      for (int t = 0; t < w.nkt; ++t) {
        const unsigned char* As = tiles + stage * STAGE_BYTES;
        const unsigned char* Bs = As + A_BYTES;
        load_frags_swz<CF>(As, Bs, off[1], a1, b1, wm, wn, g);
        mma_select<CF>(acc, a0, b0, code);
        load_frags_swz<CF>(As, Bs, off[2], a0, b0, wm, wn, g);
        mma_select<CF>(acc, a1, b1, code);
        load_frags_swz<CF>(As, Bs, off[3], a1, b1, wm, wn, g);
        mma_select<CF>(acc, a0, b0, code);
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == STAGES) { nstage = 0; nphase ^= 1u; }
        if (t + 1 < w.nkt) {
          mbar_wait(full_bar + nstage, nphase);
          load_frags_swz<CF>(tiles + nstage * STAGE_BYTES, tiles + nstage * STAGE_BYTES + A_BYTES, off[0], a0, b0, wm,
                             wn, g);
        }
        mma_select<CF>(acc, a1, b1, code);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + stage);
        stage = nstage;
        phase = nphase;
      }
    }

    // ---- epilogue: C = alpha*acc + beta*C, or raw partials into the split-K workspace
    gemm_epilogue<CF, false>(acc, p, w, wm, wn, g, q);
  }
}

// K-major operand (rows x K, pitch ld, `batch` copies `stride` apart) as a 3-D tensor map {K, rows, batch}
static int make_tmap(CUtensorMap* tm, const double* base, int rows, int K, i64 ld, i64 stride, int batch, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return 1; }
  const bool batched = batch > 1 && stride != 0;
  cuuint64_t gdim[3] = {(cuuint64_t)(K > 0 ? K : 1), (cuuint64_t)rows, (cuuint64_t)(batched ? batch : 1)};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 8, (cuuint64_t)(batched ? stride : ld * (i64)rows) * 8};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t est[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), gdim, gstr, box, est,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return 1; }
  return 0;
}

// ---- split-K of the last wave ------------------------------------------------------------------------------------
// A long-K product with U units runs floor(U / #SM) full waves and one wave with U mod #SM units that costs a whole unit
// time however few they are (8-GPU shapes: 1128 units = 7 waves + 92 units -> 8 unit times for 7.62 of work).  The full
// waves are launched as they are; the remaining output tiles get their own launch with K split s ways (s chosen so that
// rem * s fills whole waves of 1/s-units), raw partials in compact BM x BN slots of a library-owned workspace, and a
// reduction kernel that applies alpha / beta.  Deterministic: the partials of a tile are summed in split order.
__global__ void splitk_reduce_tail_kernel(const double* __restrict__ ws, int ksplit, int ntail, int tile0, int tiles,
                                          int tiles_m, int tiles_n, int nfast, int BM, int BN, int M, int N, double alpha,
                                          double beta, double* C, i64 ldc, i64 sC) {
  const int tidx = blockIdx.x;
  const int idx = tile0 + tidx;
  const int b = idx / tiles, t = idx - b * tiles;
  int tm, tn;
  if (nfast) { tn = t % tiles_n; tm = t / tiles_n; }
  else { tm = t % tiles_m; tn = t / tiles_m; }
  const int m0 = tm * BM, n0 = tn * BN;
  const i64 slot = (i64)BM * BN;
  for (int e = threadIdx.x; e < BM * BN; e += blockDim.x) {
    const int r = e / BN, c = e - r * BN;
    const int row = m0 + r, col = n0 + c;
    if (row >= M || col >= N) continue;
    double sum = 0.0;
    for (int z = 0; z < ksplit; ++z) sum += ws[((i64)z * ntail + tidx) * slot + e];
    double* dst = C + (i64)b * sC + (i64)row * ldc + col;
    *dst = (beta != 0.0) ? alpha * sum + beta * (*dst) : alpha * sum;
  }
}

static double* tail_workspace(size_t doubles) {
  static double* buf = nullptr;
  static size_t cap = 0;
  if (doubles > cap) {
    if (buf) cudaFree(buf);
    buf = nullptr;
    cap = 0;
    if (cudaMalloc(&buf, doubles * sizeof(double)) != cudaSuccess) { (void)cudaGetLastError(); buf = nullptr; return nullptr; }
    cap = doubles;
  }
  return buf;
}

template <class CF>
static int launch_tma(KParams& p, cudaStream_t st) {
  p.tiles_m = (p.M + CF::BM - 1) / CF::BM;
  p.tiles_n = (p.N + CF::BN - 1) / CF::BN;
  const i64 tiles = (i64)p.tiles_m * p.tiles_n;
  const i64 units = tiles * p.batch * p.ksplit;
  if (tiles > 2000000000LL || units > 2000000000LL) { set_error("b200cc_dgemm: too many tiles"); return 1; }
  p.tiles = (int)tiles;
  p.units = (int)units;
  p.nfast = p.tiles_n < p.tiles_m ? 1 : 0;
  TMapSet tm;
  const bool bc = p.bcoords != nullptr;
  if (make_tmap(&tm.a[0], p.A1, p.M, p.K1, p.lda1, p.sA1, bc ? p.nbA1 : p.batch, CF::BM)) return 1;
  if (make_tmap(&tm.b[0], p.B1, p.N, p.K1, p.ldb1, p.sB1, bc ? p.nbB1 : p.batch, CF::BN)) return 1;
  for (int sgm = 1; sgm < 4; ++sgm) { tm.a[sgm] = tm.a[0]; tm.b[sgm] = tm.b[0]; }
  if (p.K2 > 0) {
    if (make_tmap(&tm.a[1], p.A2, p.M, p.K2, p.lda2, p.sA2, bc ? p.nbA2 : p.batch, CF::BM)) return 1;
    if (make_tmap(&tm.b[1], p.B2, p.N, p.K2, p.ldb2, p.sB2, bc ? p.nbB2 : p.batch, CF::BN)) return 1;
  }
  if (p.K3 > 0) {
    if (make_tmap(&tm.a[2], p.A3, p.M, p.K3, p.lda3, p.sA3, bc ? p.nbA3 : p.batch, CF::BM)) return 1;
    if (make_tmap(&tm.b[2], p.B3, p.N, p.K3, p.ldb3, p.sB3, bc ? p.nbB3 : p.batch, CF::BN)) return 1;
  }
  if (p.K4 > 0) {
    if (make_tmap(&tm.a[3], p.A4, p.M, p.K4, p.lda4, p.sA4, bc ? p.nbA4 : p.batch, CF::BM)) return 1;
    if (make_tmap(&tm.b[3], p.B4, p.N, p.K4, p.ldb4, p.sB4, bc ? p.nbB4 : p.batch, CF::BN)) return 1;
  }
  constexpr int SMEM = STAGES * (CF::BM + CF::BN) * 128 + 2 * STAGES * (int)sizeof(uint64_t) + 1024;
  static bool configured = false;
  if (!configured) {
    B200CC_CUDA_OK(cudaFuncSetAttribute(dgemm_tma_kernel<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  // split-K of the last wave (see above): long-K, no caller split, plain output
  static const bool tail_on = [] { const char* e = getenv("B200CC_GEMM_TAIL"); return !(e && e[0] == '0'); }();
  const int nsm = sm_count();
  // Only where waves are countable: few of them (the unsynchronised end of a long run of waves costs as much as the split
  // saves -- measured at N = 1: ladder 195.8 -> 199.1 ms) and tiles of equal cost (M = 820 has a half-cost seventh row tile:
  // a uniform split by the caller balances those better -- measured at the N = 8 ladder shape: 26.5 vs 28.3 ms).
  // (a mildly ragged last row tile -- M = 1500: 12 tiles, the last at 3/4 cost -- keeps the waves countable)
  const int rem_rows = p.M % CF::BM;
  const int rag_frags = ((rem_rows + 7) / 8 + CF::WARPS_M - 1) / CF::WARPS_M;       // fragments per warp in the last row tile
  const bool ragged_m = p.tiles_m <= 16 && rem_rows != 0 && 10 * rag_frags < 7 * CF::MI;
  if (tail_on && p.ksplit == 1 && p.kt_total >= 512 && p.cube_nv == 0 && p.units > nsm && p.units <= 16 * nsm && !ragged_m) {
    const int rem = p.units % nsm;
    int best_s = 1;
    double best_c = 1.0;
    if (rem > 0) {
      for (int sp = 2; sp <= 16 && p.kt_total / sp >= 96; ++sp) {
        const double c = (double)(((i64)rem * sp + nsm - 1) / nsm) / sp + 0.01 * sp;   // waves of 1/sp-units (+ per-split overhead)
        if (c < best_c - 1e-9) { best_c = c; best_s = sp; }
      }
    }
    if (best_s > 1 && best_c < 0.9) {
      double* ws = tail_workspace((size_t)rem * best_s * CF::BM * CF::BN);
      if (ws) {
        KParams bulk = p;
        bulk.units = p.units - rem;
        dgemm_tma_kernel<CF><<<pick_grid(bulk), CF::NT + WS_PRODUCER_THREADS, SMEM, st>>>(bulk, tm);
        if (check_launch("dgemm_tma_kernel")) return 1;
        KParams tail = p;
        tail.tile0 = p.units - rem;
        tail.compact = 1;
        tail.ksplit = best_s;
        tail.ws = ws;
        tail.kt_per_split = (p.kt_total + best_s - 1) / best_s;
        tail.units = rem * best_s;
        dgemm_tma_kernel<CF><<<pick_grid(tail), CF::NT + WS_PRODUCER_THREADS, SMEM, st>>>(tail, tm);
        if (check_launch("dgemm_tma_kernel")) return 1;
        splitk_reduce_tail_kernel<<<rem, 256, 0, st>>>(ws, best_s, rem, tail.tile0, p.tiles, p.tiles_m, p.tiles_n, p.nfast,
                                                       CF::BM, CF::BN, p.M, p.N, p.alpha, p.beta, p.C, p.ldc, p.sC);
        return check_launch("splitk_reduce_tail_kernel");
      }
    }
  }
  const int grid = pick_grid(p);
  dgemm_tma_kernel<CF><<<grid, CF::NT + WS_PRODUCER_THREADS, SMEM, st>>>(p, tm);
  return check_launch("dgemm_tma_kernel");
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

template <class CF, bool TA, bool TB, int VEC>
static int launch(const KParams& p, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    B200CC_CUDA_OK(cudaFuncSetAttribute(dgemm_kernel<CF, TA, TB, VEC>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
    configured = true;
  }
  dgemm_kernel<CF, TA, TB, VEC><<<grid, CF::NT, CF::SMEM_BYTES, st>>>(p);
  return check_launch("dgemm_kernel");
}

template <class CF, bool WS>
static int dispatch(KParams& p, int ta, int tb, bool v2, cudaStream_t st) {
  p.tiles_m = (p.M + CF::BM - 1) / CF::BM;
  p.tiles_n = (p.N + CF::BN - 1) / CF::BN;
  const i64 tiles = (i64)p.tiles_m * p.tiles_n;
  const i64 units = tiles * p.batch * p.ksplit;
  if (tiles > 2000000000LL || units > 2000000000LL) { set_error("b200cc_dgemm: too many tiles"); return 1; }
  p.tiles = (int)tiles;
  p.units = (int)units;
  // raster: the index with FEWER tiles varies fastest, so co-running CTAs share the larger operand panel in L2
  p.nfast = p.tiles_n < p.tiles_m ? 1 : 0;
  const int grid = pick_grid(p);
#define B200CC_GO(TA, TB)                                                                         \
  do {                                                                                            \
    if constexpr (WS) return v2 ? launch_ws<CF, TA, TB, 2>(p, grid, st) : launch_ws<CF, TA, TB, 1>(p, grid, st); \
    else return v2 ? launch<CF, TA, TB, 2>(p, grid, st) : launch<CF, TA, TB, 1>(p, grid, st);    \
  } while (0)
  if (!ta && !tb) B200CC_GO(false, false);
  if (!ta && tb) B200CC_GO(false, true);
  if (ta && !tb) B200CC_GO(true, false);
  B200CC_GO(true, true);
#undef B200CC_GO
}

static inline bool even(i64 x) { return (x & 1) == 0; }
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// effective padded extent of a dimension of size n cut into tiles of `tile` rows handled as interleaved
// 8-row fragments over `warps` warps (a ragged last tile costs ceil(frags/warps)/MI of a full one)
static double eff_extent(int n, int tile, int warps) {
  const int full = n / tile, rem = n - full * tile;
  if (rem == 0) return (double)full * tile;
  const int frags = (rem + 7) / 8;
  const int per_warp = (frags + warps - 1) / warps;
  return (double)full * tile + (double)per_warp * warps * 8;
}

}  // namespace b200cc

using namespace b200cc;

extern "C" int b200cc_dgemm(const b200cc_gemm_desc* d, void* stream) {
  if (!d) { set_error("b200cc_dgemm: null descriptor"); return 1; }
  if (d->struct_size != (int)sizeof(b200cc_gemm_desc)) {
    set_error("b200cc_dgemm: descriptor is %d bytes, this library's b200cc_gemm_desc is %d (stale binding? see include/b200cc.h)",
              d->struct_size, (int)sizeof(b200cc_gemm_desc));
    return 1;
  }
  if (d->M < 0 || d->N < 0 || d->K1 < 0 || d->K2 < 0 || d->batch < 0) {
    set_error("b200cc_dgemm: negative dimension"); return 1;
  }
  if (d->M == 0 || d->N == 0 || d->batch == 0) return 0;
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
  p.kt1 = (d->K1 + BK - 1) / BK;
  p.kt_total = p.kt1 + (d->K2 + BK - 1) / BK;
  // optional third / fourth K segment (the paired (T) products): TMA kernels only, checked below
  p.K3 = d->K3 > 0 ? d->K3 : 0; p.K4 = (p.K3 > 0 && d->K4 > 0) ? d->K4 : 0;
  p.A3 = d->A3; p.B3 = d->B3; p.A4 = d->A4; p.B4 = d->B4;
  p.lda3 = d->lda3; p.ldb3 = d->ldb3; p.lda4 = d->lda4; p.ldb4 = d->ldb4;
  p.sA3 = p.K3 > 0 ? d->strideA3 : 0; p.sB3 = p.K3 > 0 ? d->strideB3 : 0;
  p.sA4 = p.K4 > 0 ? d->strideA4 : 0; p.sB4 = p.K4 > 0 ? d->strideB4 : 0;
  p.nbA3 = d->nbA3; p.nbB3 = d->nbB3; p.nbA4 = d->nbA4; p.nbB4 = d->nbB4;
  p.kt2 = p.kt_total;
  p.kt_total += (p.K3 + BK - 1) / BK;
  p.kt3 = p.kt_total;
  p.kt_total += (p.K4 + BK - 1) / BK;
  p.bc_stride = p.K3 > 0 ? 8 : 4;
  if (d->K3 > 0 && d->K2 <= 0) { set_error("b200cc_dgemm: a third K segment needs the second one"); return 1; }
  p.ksplit = ksplit; p.batch = d->batch; p.ws = d->workspace;
  p.tile0 = 0; p.compact = 0;
  p.kt_per_split = (p.kt_total + ksplit - 1) / ksplit;
  if (p.kt_per_split < 1) p.kt_per_split = 1;
  p.bcoords = d->bcoords;
  p.cube_nv = d->out_cube_nv;
  p.nbA1 = d->nbA1; p.nbB1 = d->nbB1; p.nbA2 = d->nbA2; p.nbB2 = d->nbB2;

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
  const int ta = d->transA ? 1 : 0, tb = d->transB ? 1 : 0;

  bool seg34 = true;
  if (p.K3 > 0)
    seg34 = al16(p.A3) && al16(p.B3) && even(p.lda3) && even(p.ldb3) && even(p.sA3) && even(p.sB3) &&
            (p.K4 == 0 || (al16(p.A4) && al16(p.B4) && even(p.lda4) && even(p.ldb4) && even(p.sA4) && even(p.sB4)));
  const bool tma_ok = !ta && !tb && !d->table && v2 && d->K1 > 0 && seg34;
  if (p.K3 > 0 && !tma_ok) {
    set_error("b200cc_dgemm: K3/K4 segments need the TMA path (K-major, 16-byte aligned operands and strides, no table)");
    return 1;
  }
  if (d->bcoords && !tma_ok) {
    set_error("b200cc_dgemm: bcoords needs the TMA path (K-major, 16-byte aligned operands and strides)");
    return 1;
  }
  // tile configuration: 0 = auto
  int cfg = d->config;
  if (cfg == 0) {
    // Warp-specialised kernels; pick the CTA tile by a small cost model: persistent CTAs do
    // ceil(units / #SM) rounds of the average unit cost (ragged edges counted at fragment granularity).
    // The 80x128 tile wins when M is ragged for 128 (M = o^2 = 400) or when it fills the last round better.
    const int nsm = sm_count();
    auto cost = [&](int bm, int wm, int bn, int wn, double eff) {
      const double area = eff_extent(d->M, bm, wm) * eff_extent(d->N, bn, wn);
      const double units = (double)((d->M + bm - 1) / bm) * ((d->N + bn - 1) / bn) * d->batch * ksplit;
      const double rounds = (double)((i64)((units + nsm - 1) / nsm));
      return rounds * (area * d->batch * ksplit / units) / eff;
    };
    const double c4 = cost(128, 4, 128, 2, 1.0), c5 = cost(80, 2, 128, 4, 0.97);
    cfg = (d->M >= 80 && c5 < 0.97 * c4) ? 5 : 4;
    if (tma_ok) cfg += 2;   // same tiles, operands staged by the TMA unit instead of cp.async producer warps
    // one extent <= 40 (= o at the bench shapes): a 40-wide tile wastes no fragment of it (config 8/9 = 40 x 256 /
    // 256 x 40; +2 = TMA).  kernels.auto_ksplit mirrors this rule when it counts output tiles.
    static const bool skinny = [] { const char* e = getenv("B200CC_GEMM_SKINNY"); return !(e && e[0] == '0'); }();
    if (skinny && d->out_cube_nv == 0 && p.K3 == 0 && !d->bcoords) {
      if (d->M <= 40) cfg = tma_ok ? 10 : 8;
      else if (d->N <= 40) cfg = tma_ok ? 11 : 9;
    }
  }
  if (d->out_cube_nv > 0 && (cfg != 6 && cfg != 7 || d->beta != 0.0 || ksplit > 1)) {
    set_error("b200cc_dgemm: out_cube_nv needs a TMA kernel (config 6/7), beta = 0 and no split-K");
    return 1;
  }
  if (p.K3 > 0 && cfg != 6 && cfg != 7) { set_error("b200cc_dgemm: K3/K4 segments are only implemented by the TMA kernels (config 6/7)"); return 1; }
  if (d->bcoords && cfg != 6 && cfg != 7) { set_error("b200cc_dgemm: bcoords is only implemented by the TMA kernels (config 6/7)"); return 1; }
  int rc;
  if (cfg == 2) rc = dispatch<CfgB, false>(p, ta, tb, v2, st);
  else if (cfg == 4) rc = dispatch<CfgB, true>(p, ta, tb, v2, st);
  else if (cfg == 5) rc = dispatch<CfgC, true>(p, ta, tb, v2, st);
  else if (cfg == 8) rc = dispatch<CfgD, true>(p, ta, tb, v2, st);
  else if (cfg == 9) rc = dispatch<CfgE, true>(p, ta, tb, v2, st);
  else if (cfg == 10 || cfg == 11) {
    if (!tma_ok) { set_error("b200cc_dgemm: config 10/11 (TMA) needs K-major 16-byte aligned operands without an address table"); return 1; }
    rc = cfg == 10 ? launch_tma<CfgD>(p, st) : launch_tma<CfgE>(p, st);
  }
  else if (cfg == 6) {
    if (!tma_ok) { set_error("b200cc_dgemm: config 6 (TMA) needs K-major 16-byte aligned operands without an address table"); return 1; }
    rc = launch_tma<CfgB>(p, st);
  } else if (cfg == 7) {
    if (!tma_ok) { set_error("b200cc_dgemm: config 7 (TMA) needs K-major 16-byte aligned operands without an address table"); return 1; }
    rc = launch_tma<CfgC>(p, st);
  }
  else { set_error("b200cc_dgemm: unknown tile config %d", cfg); return 1; }
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
