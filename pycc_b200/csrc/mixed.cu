// mixed.cu -- the mixed-precision contraction path (BASELINE configs[4]: "TF32 contractions, FP64 accumulation").
//
//   C[b](m,n) = alpha * sum_k A[b](m,k) B[b](n,k) + beta * C[b](m,n)          C, alpha, beta: FP64
//
// Every FP64 operand is split ONCE into two TF32-representable FP32 planes, x ~= hi + lo (22+ significant bits;
// b200cc_split_tf32), and the product is evaluated as  Ahi.Bhi + Ahi.Blo + Alo.Bhi  ("3xTF32": the dropped
// Alo.Blo term is 2^-22 relative) on the 5th-generation tensor cores:
//
//   * tcgen05.mma.cta_group::1.kind::tf32, M=128 x N=BN x K=8 per instruction, issued by ONE thread; both operands
//     come from shared memory through 64-bit matrix descriptors (K-major, 128-byte or 64-byte swizzle);
//   * operand tiles are staged by TMA (cp.async.bulk.tensor.3d, same swizzle) into an NSTAGE ring, four boxes per
//     stage (A_hi, A_lo, B_hi, B_lo), completion on the stage's mbarrier; tcgen05.commit frees the slot;
//   * the FP32 accumulator lives in TMEM (2 x BN columns: two buffers).  FP32 accumulation is only trusted for
//     `kchunk` summation indices: after each K chunk the MMA thread commits the buffer to the epilogue warps and
//     continues into the other buffer, while the four epilogue warps read the finished chunk with tcgen05.ld
//     (32 lanes x 32 columns per instruction) and add it IN FP64 to the output tile in global memory
//     (first chunk: alpha*acc + beta*C, later chunks: C += alpha*acc).  The running FP64 sum therefore never
//     occupies registers (a 128x256 FP64 tile would be the whole register file) and the tensor pipe never waits for
//     the drain;
//   * persistent CTAs (one per SM), work unit = output tile, M-tile index fastest so that concurrently running
//     CTAs share the B panel (for the ladder: the rows of <ab|ef>) in L2.
//
// Roles: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane), warps 2..5 = epilogue
// (a warp may only touch the TMEM lane quarter  warp_id % 4).
//
// Replaces, in precision='MP' mode, the same reference lines as b200cc_dgemm (ccwfn.py:931 ladder, the o^3v^3 ring
// terms 644-645/683/715/933-935, Wmnij 603) -- see include/b200cc.h.
#include <cuda.h>

#include "common.cuh"

namespace b200cc {
namespace mx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n"
      " bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, TF32 inputs, FP32 accumulate; acc == 0 overwrites D
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes (the warp's TMEM lane quarter) x 32 consecutive 32-bit columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a thread that owns a run of a row moves one full 32-byte sector
// per instruction
__device__ __forceinline__ void ld_global_v4(const double* p, double (&v)[4]) {
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];\n" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st_global_v4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// Shared-memory matrix descriptor of a K-major operand tile whose rows are ROW_BYTES (= the swizzle span) long:
// 8-row groups are 8*ROW_BYTES apart (stride byte offset); start address / offsets in 16-byte units;
// bits [46,48) = descriptor version 1 (sm_100); bits [61,64) = swizzle mode (2 = 128 B, 4 = 64 B).
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  static_assert(ROW_BYTES == 128 || ROW_BYTES == 64, "swizzle span");
  constexpr uint64_t SBO = (8 * ROW_BYTES) >> 4;
  constexpr uint64_t LT = ROW_BYTES == 128 ? 2 : 4;
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (SBO << 32) | (1ull << 46) | (LT << 61);
}

// Instruction descriptor (kind::tf32): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major
// (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
template <int BN>
__host__ __device__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

struct MxParams {
  int M, N, batch;
  int tiles_m, tiles_n, units;
  int nkb;           // k-blocks (of BK) in the summation (both segments)
  int nkb1;          // k-blocks of the first K segment (kernel R: an optional second segment follows, own operands)
  const int* bcoords;  // [batch][4] per-batch slab index of {A1, B1, A2, B2}, or nullptr (batch index / shared operand)
  int bA2, bB2;
  int kc;            // k-blocks accumulated in FP32 before a drain into FP64
  int bA, bB;        // operand has its own batch dimension (else shared by all batch entries)
  double* C;
  i64 ldc, sC;
  double alpha, beta;
  int vec;           // widest legal vector access to C in doubles: 4 (256-bit LDG/STG, one full sector per thread), 2, or 0
  int gm;            // rasterisation: units sweep the n-tiles inside bands of gm m-tiles (one wave ~ gm x 148/gm tiles)
  // leader/follower schedule (kernel R, prog != nullptr): the grid is a gm x gn block of CTAs, CTA c = (i = c % gm,
  // g = c / gm) computes tile (mb*gm + i, nb*gn + g) of block (mb, nb) in round R = (b*nmb + mb)*nnb + nb
  int* prog;         // [grid] number of k-blocks whose operands have ARRIVED in the CTA's shared memory (R*nkb + kb + 1)
  int gn, nmb, nnb, nrounds;
  int skew;          // followers poll their leaders every `skew` k-blocks and stay that far behind them
};

// unit -> (m-tile, n-tile, batch).  Within a band of gm m-tiles the m index runs fastest, so the ~148 tiles in flight
// form a gm x (148/gm) block: gm + 148/gm distinct operand panels instead of 148 + 1.
__device__ __forceinline__ void decode_unit(const MxParams& p, int u, int& tm, int& tn, int& b) {
  const int per_b = p.tiles_m * p.tiles_n;
  b = u / per_b;
  const int r = u - b * per_b;
  const int band_sz = p.gm * p.tiles_n;
  const int band = r / band_sz;
  const int m_base = band * p.gm;
  const int gsz = min(p.gm, p.tiles_m - m_base);
  const int w = r - band * band_sz;
  tn = w / gsz;
  tm = m_base + (w - tn * gsz);
}

// Leader/follower tile schedule (OPTIONAL, off by default -- kept as a measured negative result).
// The o=40,v=300 ladder reads 953 GB from DRAM for 66 GB of operands (profiles/ladder_mp_o40_r01_raw.csv): the tiles
// shared by concurrently running CTAs mostly miss L2.  This schedule makes the sharing explicit: the grid is a gm x gn
// block of CTAs, CTA (0,0) leads, CTAs (0,g) / (i,0) follow it, everybody else follows those two, and a follower only
// requests k-block kb once its leaders' copy of that block has ARRIVED in their shared memory (the leader's MMA thread
// publishes its count after the full barrier), i.e. when the lines are in L2.  Measured (profiles/mp_ncu_lf_r01.csv,
// ladder slice 1600 x 36000 x 90000): DRAM reads 203 -> 164 GB, L2 hit rate 51 -> 61 %, but the run time does not
// improve (54.0 -> 56.7 ms): the kernel is bound by the L2 -> shared-memory path (tensor pipe 59-70 % active even for
// the o^3v^3 shapes that read only 20 GB from DRAM), not by DRAM.  All CTAs are co-resident (grid <= SM count, one CTA
// per SM); a spin limit drops the wait rather than hang if that ever fails.
__device__ __forceinline__ int ld_volatile(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile(int* p, int v) {
  asm volatile("st.volatile.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_progress(const int* slot, int need, bool& on) {
  int spins = 0;
  while (on && ld_volatile(slot) < need) {
    if (++spins > (1 << 20)) on = false;
    __nanosleep(40);
  }
}

// iteration `it` of a CTA -> its tile.  Returns 0 = no more work, 1 = tile (tm, tn, b), 2 = idle in this round
// (leader/follower schedule only: ragged edge of the tile grid).
__device__ __forceinline__ int get_work(const MxParams& p, int it, int G, int& tm, int& tn, int& b) {
  if (p.prog == nullptr) {
    const int u = blockIdx.x + it * G;
    if (u >= p.units) return 0;
    decode_unit(p, u, tm, tn, b);
    return 1;
  }
  if (it >= p.nrounds) return 0;
  const int nb = it % p.nnb;
  const int r = it / p.nnb;
  const int mb = r % p.nmb;
  b = r / p.nmb;
  tm = mb * p.gm + static_cast<int>(blockIdx.x) % p.gm;
  tn = nb * p.gn + static_cast<int>(blockIdx.x) / p.gm;
  return (tm < p.tiles_m && tn < p.tiles_n) ? 1 : 2;
}

constexpr int MX_THREADS = 192;
constexpr int BM = 128;

template <int BN, int BK, int NSTAGE>
__global__ void __launch_bounds__(MX_THREADS, 1)
    tf32x3_gemm_kernel(const MxParams p, const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                       const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int ROW_BYTES = BK * 4;
  constexpr int A_BYTES = BM * ROW_BYTES, B_BYTES = BN * ROW_BYTES, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr uint32_t IDESC = make_idesc<BN>();
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + NSTAGE * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tfull_bar = empty_bar + NSTAGE;   // accumulator buffer b holds a finished chunk
  uint64_t* tempty_bar = tfull_bar + 2;       // accumulator buffer b has been drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(tfull_bar + 0, 1);
    mbar_init(tfull_bar + 1, 1);
    mbar_init(tempty_bar + 0, 4);
    mbar_init(tempty_bar + 1, 4);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const int nchunks = (p.nkb + p.kc - 1) / p.kc;

  if (warp == 0) {
    // ====================================== TMA producer ======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < p.units; u += G) {
        int tm, tn, b;
        decode_unit(p, u, tm, tn, b);
        const int m0 = tm * BM, n0 = tn * BN;
        const int ba = p.bA ? b : 0, bb = p.bB ? b : 0;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1u);
          unsigned char* s = tiles + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar + stage, STAGE_BYTES);
          tma_load_3d(s, &tmAh, kb * BK, m0, ba, full_bar + stage);
          tma_load_3d(s + 2 * A_BYTES, &tmBh, kb * BK, n0, bb, full_bar + stage);
          tma_load_3d(s + A_BYTES, &tmAl, kb * BK, m0, ba, full_bar + stage);
          tma_load_3d(s + 2 * A_BYTES + B_BYTES, &tmBl, kb * BK, n0, bb, full_bar + stage);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_it = 0;
      for (int u = blockIdx.x; u < p.units; u += G) {
        for (int c = 0; c < nchunks; ++c, ++acc_it) {
          const uint32_t buf = acc_it & 1u, aphase = (acc_it >> 1) & 1u;
          mbar_wait(tempty_bar + buf, aphase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          const int kb0 = c * p.kc, kb1 = min(p.nkb, kb0 + p.kc);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full_bar + stage, phase);
            tc_fence_after();
            const uint32_t s = smem_u32(tiles + stage * STAGE_BYTES);
            const uint64_t ah = make_sdesc<ROW_BYTES>(s), al = make_sdesc<ROW_BYTES>(s + A_BYTES);
            const uint64_t bh = make_sdesc<ROW_BYTES>(s + 2 * A_BYTES), bl = make_sdesc<ROW_BYTES>(s + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int j = 0; j < BK / 8; ++j) {
              // one k-step = 8 TF32 = 32 bytes along the row: +2 in the descriptor's 16-byte address units
              const uint64_t o = static_cast<uint64_t>(2 * j);
              umma_tf32(d_tmem, ah + o, bh + o, IDESC, (kb > kb0 || j > 0) ? 1u : 0u);
              umma_tf32(d_tmem, ah + o, bl + o, IDESC, 1u);
              umma_tf32(d_tmem, al + o, bh + o, IDESC, 1u);
            }
            umma_commit(empty_bar + stage);            // slot reusable once these MMAs have read it
            if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
          }
          umma_commit(tfull_bar + buf);                // chunk complete -> epilogue
        }
      }
    }
  } else {
    // ====================================== epilogue: TMEM -> FP64 global ======================================
    const int quarter = warp & 3;
    const int rloc = quarter * 32 + lane;
    uint32_t acc_it = 0;
    for (int u = blockIdx.x; u < p.units; u += G) {
      int tm, tn, b;
      decode_unit(p, u, tm, tn, b);
      const int row = tm * BM + rloc;
      const int n0 = tn * BN;
      double* crow = p.C + (i64)b * p.sC + (i64)row * p.ldc + n0;
      for (int c = 0; c < nchunks; ++c, ++acc_it) {
        const uint32_t buf = acc_it & 1u, aphase = (acc_it >> 1) & 1u;
        mbar_wait(tfull_bar + buf, aphase);
        tc_fence_after();
        const double fb = c == 0 ? p.beta : 1.0;       // weight of what is already in C
        const bool rd = fb != 0.0;
        const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * BN;
#pragma unroll 1
        for (int cb = 0; cb < BN / 32; ++cb) {
          const int col0 = n0 + cb * 32;
          if (col0 >= p.N) break;                      // warp-uniform
          uint32_t v[32];
          tmem_ld32(tbase + cb * 32, v);
          tmem_wait_ld();
          if (row < p.M) {
            double* cp = crow + cb * 32;
            if (p.vec && col0 + 32 <= p.N) {
              double2 o[16];
              if (rd) {
#pragma unroll
                for (int t = 0; t < 16; ++t) o[t] = *reinterpret_cast<const double2*>(cp + 2 * t);
              }
#pragma unroll
              for (int t = 0; t < 16; ++t) {
                double x = p.alpha * static_cast<double>(__uint_as_float(v[2 * t]));
                double y = p.alpha * static_cast<double>(__uint_as_float(v[2 * t + 1]));
                if (rd) { x += fb * o[t].x; y += fb * o[t].y; }
                *reinterpret_cast<double2*>(cp + 2 * t) = make_double2(x, y);
              }
            } else {
#pragma unroll
              for (int t = 0; t < 32; ++t) {
                if (col0 + t < p.N) {
                  double x = p.alpha * static_cast<double>(__uint_as_float(v[t]));
                  if (rd) x += fb * cp[t];
                  cp[t] = x;
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + buf);
      }
    }
  }

  // ---- teardown: every MMA has completed (the epilogue waited for the last chunk) and TMEM has been read
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// =====================================================================================================
// Kernel R ("register accumulation") -- the default for long summations.
//
// Measured on B200 (profiles/mp_probe_r01_*.json): the FP32 accumulate of tcgen05.mma TRUNCATES (round toward
// zero), so an accumulator that receives n MMAs comes out scaled by (1 - n * 1.6e-8): a systematic shrink, not noise
// (kchunk = 2048 in the kernel above: 768 MMAs per chunk, bias -1.25e-5).  A 1e-6 Eh energy needs that bias below
// ~3e-7, i.e. FP32 runs of a few dozen MMAs.  Hence here:
//   * the Ahi.Bhi products go to their own accumulator; the two cross products (2^-11 smaller, their truncation is
//     irrelevant) to a second one -- the large accumulator sees ONE MMA per k-step instead of three;
//   * chunks are short (default 256 summation indices = 32 MMAs on the large accumulator: bias 5e-7 of the chunk sum);
//   * the running FP64 sum lives in the REGISTERS of 8 epilogue warps (128x128 tile: 64 doubles per thread), so a drain
//     is TMEM -> registers only (no global traffic) and can be afforded every 256 k: per chunk and thread
//     8 tcgen05.ld.x16, 64 FADD (hh + cross, round-to-nearest), 64 F2F, 64 DADD, hidden behind the next chunk's MMAs
//     (both accumulators are double-buffered: 4 x 128 = 512 TMEM columns).
// =====================================================================================================
constexpr int MXR_THREADS = 320;   // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue
constexpr int RBN = 128;

template <int BK, int NSTAGE>
__global__ void __launch_bounds__(MXR_THREADS, 1)
    tf32x3_gemm_r_kernel(const MxParams p, const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                         const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                         const __grid_constant__ CUtensorMap tmA2h, const __grid_constant__ CUtensorMap tmA2l,
                         const __grid_constant__ CUtensorMap tmB2h, const __grid_constant__ CUtensorMap tmB2l) {

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int BN = RBN;
  constexpr int ROW_BYTES = BK * 4;
  constexpr int A_BYTES = BM * ROW_BYTES, B_BYTES = BN * ROW_BYTES, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = 512;          // hh[0], hh[1], cross[0], cross[1], 128 columns each
  constexpr uint32_t IDESC = make_idesc<BN>();
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + NSTAGE * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tfull_bar = empty_bar + NSTAGE;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(tfull_bar + 0, 1);
    mbar_init(tfull_bar + 1, 1);
    mbar_init(tempty_bar + 0, 8);
    mbar_init(tempty_bar + 1, 8);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const int nchunks = (p.nkb + p.kc - 1) / p.kc;

  if (warp == 0) {
    // ---- TMA producer (one lane); in the leader/follower schedule it first waits for its leaders' data to be in L2
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const bool lf = p.prog != nullptr;
      const int ci = static_cast<int>(blockIdx.x) % p.gm, cg = static_cast<int>(blockIdx.x) / p.gm;
      const int* leadB = p.prog + cg * p.gm;      // CTA (0, g): same n-tile, first to fetch the B tile
      const int* leadA = p.prog + ci;             // CTA (i, 0): same m-tile, first to fetch the A tile
      bool on = lf;
      for (int it = 0;; ++it) {
        int tm, tn, b;
        const int w = get_work(p, it, G, tm, tn, b);
        if (w == 0) break;
        if (w == 2) continue;
        const int m0 = tm * BM, n0 = tn * BN;
        int ba = p.bA ? b : 0, bb = p.bB ? b : 0, ba2 = p.bA2 ? b : 0, bb2 = p.bB2 ? b : 0;
        if (p.bcoords) {     // (T): each batch entry names its own slab of every operand
          const int4 co = reinterpret_cast<const int4*>(p.bcoords)[b];
          ba = co.x; bb = co.y; ba2 = co.z; bb2 = co.w;
        }
        for (int kb = 0; kb < p.nkb; ++kb) {
          if (lf && kb % p.skew == 0) {
            const int need = it * p.nkb + min(kb + p.skew, p.nkb);
            if (ci != 0) wait_progress(leadB, need, on);
            if (cg != 0) wait_progress(leadA, need, on);
          }
          mbar_wait(empty_bar + stage, phase ^ 1u);
          unsigned char* s = tiles + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar + stage, STAGE_BYTES);
          if (kb < p.nkb1) {
            tma_load_3d(s, &tmAh, kb * BK, m0, ba, full_bar + stage);
            tma_load_3d(s + 2 * A_BYTES, &tmBh, kb * BK, n0, bb, full_bar + stage);
            tma_load_3d(s + A_BYTES, &tmAl, kb * BK, m0, ba, full_bar + stage);
            tma_load_3d(s + 2 * A_BYTES + B_BYTES, &tmBl, kb * BK, n0, bb, full_bar + stage);
          } else {           // second K segment (its own operands), accumulated into the same tile
            const int k2 = (kb - p.nkb1) * BK;
            tma_load_3d(s, &tmA2h, k2, m0, ba2, full_bar + stage);
            tma_load_3d(s + 2 * A_BYTES, &tmB2h, k2, n0, bb2, full_bar + stage);
            tma_load_3d(s + A_BYTES, &tmA2l, k2, m0, ba2, full_bar + stage);
            tma_load_3d(s + 2 * A_BYTES + B_BYTES, &tmB2l, k2, n0, bb2, full_bar + stage);
          }
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_it = 0;
      for (int it = 0;; ++it) {
        int tm_, tn_, b_;
        const int w = get_work(p, it, G, tm_, tn_, b_);
        if (w == 0) break;
        if (w == 2) {                       // idle round: its followers must not wait for this CTA
          st_volatile(p.prog + blockIdx.x, (it + 1) * p.nkb);
          continue;
        }
        for (int c = 0; c < nchunks; ++c, ++acc_it) {
          const uint32_t buf = acc_it & 1u, aphase = (acc_it >> 1) & 1u;
          mbar_wait(tempty_bar + buf, aphase ^ 1u);
          tc_fence_after();
          const uint32_t d_hh = tmem_base + buf * BN, d_x = tmem_base + 2 * BN + buf * BN;
          const int kb0 = c * p.kc, kb1 = min(p.nkb, kb0 + p.kc);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full_bar + stage, phase);
            if (p.prog != nullptr && ((kb + 1) % p.skew == 0 || kb + 1 == p.nkb))
              st_volatile(p.prog + blockIdx.x, it * p.nkb + kb + 1);     // these operands are now in L2 for the followers
            tc_fence_after();
            const uint32_t s = smem_u32(tiles + stage * STAGE_BYTES);
            const uint64_t ah = make_sdesc<ROW_BYTES>(s), al = make_sdesc<ROW_BYTES>(s + A_BYTES);
            const uint64_t bh = make_sdesc<ROW_BYTES>(s + 2 * A_BYTES), bl = make_sdesc<ROW_BYTES>(s + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int j = 0; j < BK / 8; ++j) {
              const uint64_t o = static_cast<uint64_t>(2 * j);
              const uint32_t cont = (kb > kb0 || j > 0) ? 1u : 0u;
              umma_tf32(d_hh, ah + o, bh + o, IDESC, cont);
              umma_tf32(d_x, ah + o, bl + o, IDESC, cont);
              umma_tf32(d_x, al + o, bh + o, IDESC, 1u);
            }
            umma_commit(empty_bar + stage);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
          }
          umma_commit(tfull_bar + buf);
        }
      }
    }
  } else {
    // ---- epilogue: warp w owns TMEM lanes 32*(w%4).. and columns 64*half..; 64 FP64 running sums per thread
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int rloc = quarter * 32 + lane;
    uint32_t acc_it = 0;
    double acc[64];
    for (int it = 0;; ++it) {
      int tm, tn, b;
      const int w = get_work(p, it, G, tm, tn, b);
      if (w == 0) break;
      if (w == 2) continue;
      const int row = tm * BM + rloc;
      const int n0 = tn * BN + half * 64;
#pragma unroll
      for (int t = 0; t < 64; ++t) acc[t] = 0.0;
      for (int c = 0; c < nchunks; ++c, ++acc_it) {
        const uint32_t buf = acc_it & 1u, aphase = (acc_it >> 1) & 1u;
        mbar_wait(tfull_bar + buf, aphase);
        tc_fence_after();
        const uint32_t t_hh = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * BN + half * 64;
        const uint32_t t_x = t_hh + 2 * BN;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t v[16], w[16];
          tmem_ld16(t_hh + q * 16, v);
          tmem_ld16(t_x + q * 16, w);
          tmem_wait_ld();
#pragma unroll
          for (int t = 0; t < 16; ++t)
            acc[q * 16 + t] += static_cast<double>(__uint_as_float(v[t]) + __uint_as_float(w[t]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + buf);
      }
      // ---- tile finished: C = alpha * acc + beta * C
      if (row < p.M && n0 < p.N) {
        double* cp = p.C + (i64)b * p.sC + (i64)row * p.ldc + n0;
        const bool rd = p.beta != 0.0;
        if (p.vec == 4 && n0 + 64 <= p.N) {
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            double o[4] = {0.0, 0.0, 0.0, 0.0};
            if (rd) ld_global_v4(cp + 4 * t, o);
            st_global_v4(cp + 4 * t, p.alpha * acc[4 * t] + p.beta * o[0], p.alpha * acc[4 * t + 1] + p.beta * o[1],
                         p.alpha * acc[4 * t + 2] + p.beta * o[2], p.alpha * acc[4 * t + 3] + p.beta * o[3]);
          }
        } else if (p.vec && n0 + 64 <= p.N) {
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            double x = p.alpha * acc[2 * t], y = p.alpha * acc[2 * t + 1];
            if (rd) {
              const double2 o = *reinterpret_cast<const double2*>(cp + 2 * t);
              x += p.beta * o.x;
              y += p.beta * o.y;
            }
            *reinterpret_cast<double2*>(cp + 2 * t) = make_double2(x, y);
          }
        } else {
#pragma unroll
          for (int t = 0; t < 64; ++t) {
            if (n0 + t < p.N) {
              double x = p.alpha * acc[t];
              if (rd) x += p.beta * cp[t];
              cp[t] = x;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =====================================================================================================
// Kernel R2: kernel R on CTA PAIRS (tcgen05 cta_group::2).  The kernel above is bound by operand delivery
// (L2 -> shared memory: a 128x128 tile needs 64 KB per 32-k stage, tensor pipe 59-70 % active).  Two CTAs of a
// cluster compute ONE 256x128 tile: each loads its own 128 rows of A but only HALF (64 rows) of B, so a stage is 48 KB
// per CTA (25 % less traffic per MMA, and 4 stages fit instead of 3).  The leader CTA's MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256: rows 0-127 accumulate in the leader's TMEM, 128-255 in the peer's; B is read from
// both CTAs' shared memory); both CTAs' TMA loads signal the LEADER's full barrier (.cta_group::2 loads, barrier
// address with the peer bit cleared); tcgen05.commit multicasts the "slot free" / "chunk finished" arrivals to both
// CTAs; each CTA's epilogue warps drain their own TMEM half and arrive remotely on the leader's "buffer drained"
// barrier (mapa + mbarrier.arrive.shared::cluster).
// =====================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the pair's even CTA
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(leader_bar) & PEER_BIT_MASK)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(smem_u32(bar)), "r"(0));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(ra) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_2sm_256x128() {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(128 >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
}

template <int BK, int NSTAGE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MXR_THREADS, 1)
    tf32x3_gemm_r2_kernel(const MxParams p, const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                          const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int BN = RBN, BNH = RBN / 2;
  constexpr int ROW_BYTES = BK * 4;
  constexpr int A_BYTES = BM * ROW_BYTES, B_BYTES = BNH * ROW_BYTES, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t IDESC = make_idesc_2sm_256x128();
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + NSTAGE * STAGE_BYTES);   // used in the leader CTA only
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tfull_bar = empty_bar + NSTAGE;
  uint64_t* tempty_bar = tfull_bar + 2;                                              // used in the leader CTA only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int pair = static_cast<int>(blockIdx.x) >> 1, npairs = static_cast<int>(gridDim.x) >> 1;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(tfull_bar + 0, 1);
    mbar_init(tfull_bar + 1, 1);
    mbar_init(tempty_bar + 0, 16);      // 8 epilogue warps x 2 CTAs
    mbar_init(tempty_bar + 1, 16);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const int nchunks = (p.nkb + p.kc - 1) / p.kc;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = pair; u < p.units; u += npairs) {
        int tm, tn, b;
        decode_unit(p, u, tm, tn, b);
        const int m0 = tm * 2 * BM + rank * BM, n0 = tn * BN + rank * BNH;
        const int ba = p.bA ? b : 0, bb = p.bB ? b : 0;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1u);
          unsigned char* s = tiles + stage * STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(full_bar + stage, 2 * STAGE_BYTES);      // both CTAs' bytes land on this barrier
          tma_load_3d_2sm(s, &tmAh, kb * BK, m0, ba, full_bar + stage);
          tma_load_3d_2sm(s + 2 * A_BYTES, &tmBh, kb * BK, n0, bb, full_bar + stage);
          tma_load_3d_2sm(s + A_BYTES, &tmAl, kb * BK, m0, ba, full_bar + stage);
          tma_load_3d_2sm(s + 2 * A_BYTES + B_BYTES, &tmBl, kb * BK, n0, bb, full_bar + stage);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_it = 0;
      for (int u = pair; u < p.units; u += npairs) {
        for (int c = 0; c < nchunks; ++c, ++acc_it) {
          const uint32_t buf = acc_it & 1u, aphase = (acc_it >> 1) & 1u;
          mbar_wait(tempty_bar + buf, aphase ^ 1u);
          tc_fence_after();
          const uint32_t d_hh = tmem_base + buf * BN, d_x = tmem_base + 2 * BN + buf * BN;
          const int kb0 = c * p.kc, kb1 = min(p.nkb, kb0 + p.kc);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full_bar + stage, phase);
            tc_fence_after();
            const uint32_t s = smem_u32(tiles + stage * STAGE_BYTES);
            const uint64_t ah = make_sdesc<ROW_BYTES>(s), al = make_sdesc<ROW_BYTES>(s + A_BYTES);
            const uint64_t bh = make_sdesc<ROW_BYTES>(s + 2 * A_BYTES), bl = make_sdesc<ROW_BYTES>(s + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int j = 0; j < BK / 8; ++j) {
              const uint64_t o = static_cast<uint64_t>(2 * j);
              const uint32_t cont = (kb > kb0 || j > 0) ? 1u : 0u;
              umma_tf32_2sm(d_hh, ah + o, bh + o, IDESC, cont);
              umma_tf32_2sm(d_x, ah + o, bl + o, IDESC, cont);
              umma_tf32_2sm(d_x, al + o, bh + o, IDESC, 1u);
            }
            umma_commit_2sm(empty_bar + stage);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
          }
          umma_commit_2sm(tfull_bar + buf);
        }
      }
    }
  } else {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int rloc = quarter * 32 + lane;
    uint32_t acc_it = 0;
    double acc[64];
    for (int u = pair; u < p.units; u += npairs) {
      int tm, tn, b;
      decode_unit(p, u, tm, tn, b);
      const int row = tm * 2 * BM + rank * BM + rloc;
      const int n0 = tn * BN + half * 64;
#pragma unroll
      for (int t = 0; t < 64; ++t) acc[t] = 0.0;
      for (int c = 0; c < nchunks; ++c, ++acc_it) {
        const uint32_t buf = acc_it & 1u, aphase = (acc_it >> 1) & 1u;
        mbar_wait(tfull_bar + buf, aphase);
        tc_fence_after();
        const uint32_t t_hh = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * BN + half * 64;
        const uint32_t t_x = t_hh + 2 * BN;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t v[16], w[16];
          tmem_ld16(t_hh + q * 16, v);
          tmem_ld16(t_x + q * 16, w);
          tmem_wait_ld();
#pragma unroll
          for (int t = 0; t < 16; ++t)
            acc[q * 16 + t] += static_cast<double>(__uint_as_float(v[t]) + __uint_as_float(w[t]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tempty_bar + buf);
      }
      if (row < p.M && n0 < p.N) {
        double* cp = p.C + (i64)b * p.sC + (i64)row * p.ldc + n0;
        const bool rd = p.beta != 0.0;
        if (p.vec == 4 && n0 + 64 <= p.N) {
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            double o[4] = {0.0, 0.0, 0.0, 0.0};
            if (rd) ld_global_v4(cp + 4 * t, o);
            st_global_v4(cp + 4 * t, p.alpha * acc[4 * t] + p.beta * o[0], p.alpha * acc[4 * t + 1] + p.beta * o[1],
                         p.alpha * acc[4 * t + 2] + p.beta * o[2], p.alpha * acc[4 * t + 3] + p.beta * o[3]);
          }
        } else if (p.vec && n0 + 64 <= p.N) {
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            double x = p.alpha * acc[2 * t], y = p.alpha * acc[2 * t + 1];
            if (rd) {
              const double2 o = *reinterpret_cast<const double2*>(cp + 2 * t);
              x += p.beta * o.x;
              y += p.beta * o.y;
            }
            *reinterpret_cast<double2*>(cp + 2 * t) = make_double2(x, y);
          }
        } else {
#pragma unroll
          for (int t = 0; t < 64; ++t) {
            if (n0 + t < p.N) {
              double x = p.alpha * acc[t];
              if (rd) x += p.beta * cp[t];
              cp[t] = x;
            }
          }
        }
      }
    }
  }

  // the peer's shared memory / TMEM must stay alive until the leader's last MMA has completed: both CTAs' epilogues
  // have seen the last "chunk finished" arrival before they get here
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// ---- FP64 -> (hi, lo) TF32 planes --------------------------------------------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// planes are [batch][rows][Kp] (Kp % 4 == 0, columns K..Kp zero); source rows x K with pitch ld, batches `stride` apart
__global__ void __launch_bounds__(256) split_tf32_kernel(const double* __restrict__ src, i64 ld, i64 stride, int rows,
                                                         int K, int Kp, int batch, float* __restrict__ hi,
                                                         float* __restrict__ lo) {
  const i64 q4 = Kp >> 2;
  const i64 total = (i64)batch * rows * q4;
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    const i64 rowg = e / q4;
    const int k0 = static_cast<int>(e - rowg * q4) * 4;
    const i64 b = rowg / rows;
    const i64 r = rowg - b * rows;
    const double* s = src + b * stride + r * ld + k0;
    float h[4], l[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const double x = (k0 + t < K) ? __ldg(s + t) : 0.0;
      h[t] = tf32_rna(static_cast<float>(x));
      l[t] = tf32_rna(static_cast<float>(x - static_cast<double>(h[t])));
    }
    *reinterpret_cast<float4*>(hi + rowg * Kp + k0) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo + rowg * Kp + k0) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// inverse of split_tf32_kernel: dst (rows x K doubles, pitch ld) = hi + lo, planes with pitch Kp (Kp % 4 == 0)
__global__ void __launch_bounds__(256) merge_tf32_kernel(const float* __restrict__ hi, const float* __restrict__ lo,
                                                         int Kp, i64 rows, int K, double* __restrict__ dst, i64 ld) {
  const i64 q4 = Kp >> 2;
  const i64 total = rows * q4;
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
    const i64 r = e / q4;
    const int k0 = static_cast<int>(e - r * q4) * 4;
    const float4 h = __ldg(reinterpret_cast<const float4*>(hi + r * Kp + k0));
    const float4 l = __ldg(reinterpret_cast<const float4*>(lo + r * Kp + k0));
    const double v[4] = {(double)h.x + (double)l.x, (double)h.y + (double)l.y, (double)h.z + (double)l.z,
                         (double)h.w + (double)l.w};
    double* d = dst + r * ld + k0;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (k0 + t < K) d[t] = v[t];
  }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// FP32 plane (rows x K, pitch ld floats, `batch` copies `stride` floats apart) as a 3-D tensor map {K, rows, batch};
// rows / k outside the extents are zero-filled by the TMA unit
static int make_tmap(CUtensorMap* tm, const float* base, int rows, int K, i64 ld, i64 stride, int batch, int box_k,
                     int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return 1; }
  const bool batched = batch > 1 && stride != 0;
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(batched ? batch : 1)};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(batched ? stride : ld * (i64)rows) * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
  cuuint32_t est[3] = {1, 1, 1};
  const CUtensorMapSwizzle swz = box_k * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, est,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (f32) failed (%d)", (int)r); return 1; }
  return 0;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// progress array of the leader/follower schedule: a ring of 16 zeroed 256-int slots per device, one per launch
static int* progress_slot(cudaStream_t st) {
  static int* ring[64] = {nullptr};
  static unsigned next[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_error("progress_slot: bad device"); return nullptr; }
  if (!ring[dev] && cudaMalloc(&ring[dev], 16 * 256 * sizeof(int)) != cudaSuccess) {
    set_error("progress_slot: cudaMalloc failed");
    return nullptr;
  }
  int* slot = ring[dev] + 256 * (next[dev]++ & 15u);
  if (cudaMemsetAsync(slot, 0, 256 * sizeof(int), st) != cudaSuccess) { set_error("progress_slot: memset failed"); return nullptr; }
  return slot;
}

template <int BN, int BK, int NSTAGE, bool REG>
static int launch(const b200cc_gemm3_desc* d, cudaStream_t st) {
  MxParams p;
  p.M = d->M; p.N = d->N; p.batch = d->batch;
  p.tiles_m = (d->M + BM - 1) / BM;
  p.tiles_n = (d->N + BN - 1) / BN;
  const i64 units = (i64)p.tiles_m * p.tiles_n * d->batch;
  if (units > 2000000000LL) { set_error("b200cc_gemm_tf32x3: too many tiles"); return 1; }
  p.units = (int)units;
  p.nkb1 = (d->K + BK - 1) / BK;
  p.nkb = p.nkb1 + (d->K2 > 0 ? (d->K2 + BK - 1) / BK : 0);
  p.bcoords = d->bcoords;
  p.bA2 = (d->batch > 1 && d->strideA2 != 0) ? 1 : 0;
  p.bB2 = (d->batch > 1 && d->strideB2 != 0) ? 1 : 0;
  const int kchunk = d->kchunk > 0 ? d->kchunk : (REG ? 256 : 2048);
  p.kc = kchunk / BK > 0 ? kchunk / BK : 1;
  p.bA = (d->batch > 1 && d->strideA != 0) ? 1 : 0;
  p.bB = (d->batch > 1 && d->strideB != 0) ? 1 : 0;
  p.C = d->C; p.ldc = d->ldc; p.sC = d->strideC;
  p.alpha = d->alpha; p.beta = d->beta;
  p.vec = (al16(d->C) && (d->ldc & 1) == 0 && (d->strideC & 1) == 0) ? 2 : 0;
  if (p.vec && (reinterpret_cast<uintptr_t>(d->C) & 31) == 0 && (d->ldc & 3) == 0 && (d->strideC & 3) == 0) p.vec = 4;
  p.gm = p.tiles_m <= 16 ? p.tiles_m : 12;
  p.prog = nullptr;
  p.gn = p.nmb = p.nnb = p.nrounds = 0;
  p.skew = 4;

  CUtensorMap tAh, tAl, tBh, tBl;
  const bool bc = d->bcoords != nullptr;
  if (make_tmap(&tAh, d->Ahi, d->M, d->K, d->lda, d->strideA, bc ? d->nbA1 : d->batch, BK, BM)) return 1;
  if (make_tmap(&tAl, d->Alo, d->M, d->K, d->lda, d->strideA, bc ? d->nbA1 : d->batch, BK, BM)) return 1;
  if (make_tmap(&tBh, d->Bhi, d->N, d->K, d->ldb, d->strideB, bc ? d->nbB1 : d->batch, BK, BN)) return 1;
  if (make_tmap(&tBl, d->Blo, d->N, d->K, d->ldb, d->strideB, bc ? d->nbB1 : d->batch, BK, BN)) return 1;
  CUtensorMap tA2h = tAh, tA2l = tAl, tB2h = tBh, tB2l = tBl;
  if (d->K2 > 0) {
    if (!REG) { set_error("b200cc_gemm_tf32x3: a second K segment needs config 0/5/6"); return 1; }
    if (make_tmap(&tA2h, d->A2hi, d->M, d->K2, d->lda2, d->strideA2, bc ? d->nbA2 : d->batch, BK, BM)) return 1;
    if (make_tmap(&tA2l, d->A2lo, d->M, d->K2, d->lda2, d->strideA2, bc ? d->nbA2 : d->batch, BK, BM)) return 1;
    if (make_tmap(&tB2h, d->B2hi, d->N, d->K2, d->ldb2, d->strideB2, bc ? d->nbB2 : d->batch, BK, BN)) return 1;
    if (make_tmap(&tB2l, d->B2lo, d->N, d->K2, d->ldb2, d->strideB2, bc ? d->nbB2 : d->batch, BK, BN)) return 1;
  }
  if (bc) {
    if (!REG) { set_error("b200cc_gemm_tf32x3: bcoords need config 0/5/6"); return 1; }
    p.bA = p.bB = p.bA2 = p.bB2 = 1;
  }
  constexpr int SMEM = NSTAGE * (2 * BM + 2 * BN) * BK * 4 + (2 * NSTAGE + 4) * (int)sizeof(uint64_t) + 16 + 1024;
  static bool configured = false;
  const int nsm = sm_count();
  int grid = p.units < nsm ? p.units : nsm;
  if constexpr (REG) {
    static_assert(BN == RBN, "kernel R tile");
    // leader/follower schedule (see above): only on request
    const int gn = (nsm / p.gm) < p.tiles_n ? (nsm / p.gm) : p.tiles_n;
    const i64 nmb = (p.tiles_m + p.gm - 1) / p.gm, nnb = gn > 0 ? (p.tiles_n + gn - 1) / gn : 0;
    const i64 nrounds = nmb * nnb * d->batch;
    const bool want = d->lockstep > 0;
    if (want && gn >= 2 && p.gm * gn <= 256 && p.nkb >= 16 && nrounds * p.nkb < 2000000000LL) {
      p.prog = progress_slot(st);
      if (!p.prog) return 1;
      p.gn = gn; p.nmb = (int)nmb; p.nnb = (int)nnb; p.nrounds = (int)nrounds;
      p.skew = d->lockstep;
      grid = p.gm * gn;
    }
    if (!configured) {
      B200CC_CUDA_OK(cudaFuncSetAttribute(tf32x3_gemm_r_kernel<BK, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      configured = true;
    }
    tf32x3_gemm_r_kernel<BK, NSTAGE><<<grid, MXR_THREADS, SMEM, st>>>(p, tAh, tAl, tBh, tBl, tA2h, tA2l, tB2h, tB2l);
    return check_launch("tf32x3_gemm_r_kernel");
  } else {
    if (!configured) {
      B200CC_CUDA_OK(cudaFuncSetAttribute(tf32x3_gemm_kernel<BN, BK, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      configured = true;
    }
    tf32x3_gemm_kernel<BN, BK, NSTAGE><<<grid, MX_THREADS, SMEM, st>>>(p, tAh, tAl, tBh, tBl);
    return check_launch("tf32x3_gemm_kernel");
  }
}

template <int BK, int NSTAGE>
static int launch_r2(const b200cc_gemm3_desc* d, cudaStream_t st) {
  MxParams p;
  p.M = d->M; p.N = d->N; p.batch = d->batch;
  if (d->K2 > 0 || d->bcoords) { set_error("b200cc_gemm_tf32x3: config 7 takes one K segment and strided batches only"); return 1; }
  p.bcoords = nullptr; p.bA2 = p.bB2 = 0;
  p.tiles_m = (d->M + 2 * BM - 1) / (2 * BM);          // 256-row pair tiles
  p.tiles_n = (d->N + RBN - 1) / RBN;
  const i64 units = (i64)p.tiles_m * p.tiles_n * d->batch;
  if (units > 2000000000LL) { set_error("b200cc_gemm_tf32x3: too many tiles"); return 1; }
  p.units = (int)units;
  p.nkb = p.nkb1 = (d->K + BK - 1) / BK;
  const int kchunk = d->kchunk > 0 ? d->kchunk : 256;
  p.kc = kchunk / BK > 0 ? kchunk / BK : 1;
  p.bA = (d->batch > 1 && d->strideA != 0) ? 1 : 0;
  p.bB = (d->batch > 1 && d->strideB != 0) ? 1 : 0;
  p.C = d->C; p.ldc = d->ldc; p.sC = d->strideC;
  p.alpha = d->alpha; p.beta = d->beta;
  p.vec = (al16(d->C) && (d->ldc & 1) == 0 && (d->strideC & 1) == 0) ? 2 : 0;
  if (p.vec && (reinterpret_cast<uintptr_t>(d->C) & 31) == 0 && (d->ldc & 3) == 0 && (d->strideC & 3) == 0) p.vec = 4;
  p.gm = p.tiles_m <= 10 ? p.tiles_m : 8;
  p.prog = nullptr;
  p.gn = p.nmb = p.nnb = p.nrounds = 0;
  p.skew = 4;
  CUtensorMap tAh, tAl, tBh, tBl;
  if (make_tmap(&tAh, d->Ahi, d->M, d->K, d->lda, d->strideA, d->batch, BK, BM)) return 1;
  if (make_tmap(&tAl, d->Alo, d->M, d->K, d->lda, d->strideA, d->batch, BK, BM)) return 1;
  if (make_tmap(&tBh, d->Bhi, d->N, d->K, d->ldb, d->strideB, d->batch, BK, RBN / 2)) return 1;
  if (make_tmap(&tBl, d->Blo, d->N, d->K, d->ldb, d->strideB, d->batch, BK, RBN / 2)) return 1;
  constexpr int SMEM = NSTAGE * (2 * BM + RBN) * BK * 4 + (2 * NSTAGE + 4) * (int)sizeof(uint64_t) + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    B200CC_CUDA_OK(cudaFuncSetAttribute(tf32x3_gemm_r2_kernel<BK, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  const int maxpairs = sm_count() / 2;
  const int npairs = p.units < maxpairs ? p.units : maxpairs;
  tf32x3_gemm_r2_kernel<BK, NSTAGE><<<2 * npairs, MXR_THREADS, SMEM, st>>>(p, tAh, tAl, tBh, tBl);
  return check_launch("tf32x3_gemm_r2_kernel");
}

}  // namespace mx
}  // namespace b200cc

using namespace b200cc;

extern "C" int b200cc_split_tf32(const double* src, b200cc_i64 ld, b200cc_i64 stride, int rows, int K, int batch,
                                 float* hi, float* lo, b200cc_i64 ldp, void* stream) {
  if (rows <= 0 || K <= 0 || batch <= 0) return 0;
  if (ldp < K || (ldp & 3) != 0) { set_error("b200cc_split_tf32: plane pitch must be a multiple of 4 and >= K"); return 1; }
  if (!src || !hi || !lo || (reinterpret_cast<uintptr_t>(hi) & 15) || (reinterpret_cast<uintptr_t>(lo) & 15)) {
    set_error("b200cc_split_tf32: null or misaligned plane");
    return 1;
  }
  if (ldp > 2147483647LL) { set_error("b200cc_split_tf32: pitch too large"); return 1; }
  const i64 total = (i64)batch * rows * (ldp >> 2);
  i64 blocks = (total + 255) / 256;
  const i64 cap = (i64)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  mx::split_tf32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, ld, stride, rows, K, (int)ldp, batch, hi, lo);
  return check_launch("split_tf32_kernel");
}

extern "C" int b200cc_merge_tf32(const float* hi, const float* lo, b200cc_i64 ldp, b200cc_i64 rows, int K, double* dst,
                                 b200cc_i64 ld, void* stream) {
  if (rows <= 0 || K <= 0) return 0;
  if (ldp < K || (ldp & 3) != 0 || ldp > 2147483647LL || ld < K) {
    set_error("b200cc_merge_tf32: plane pitch must be a multiple of 4 and >= K, ld >= K");
    return 1;
  }
  if (!dst || !hi || !lo || (reinterpret_cast<uintptr_t>(hi) & 15) || (reinterpret_cast<uintptr_t>(lo) & 15)) {
    set_error("b200cc_merge_tf32: null or misaligned plane");
    return 1;
  }
  const i64 total = rows * (ldp >> 2);
  i64 blocks = (total + 255) / 256;
  const i64 cap = (i64)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  mx::merge_tf32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(hi, lo, (int)ldp, rows, K, dst, ld);
  return check_launch("merge_tf32_kernel");
}

extern "C" int b200cc_gemm_tf32x3(const b200cc_gemm3_desc* d, void* stream) {
  if (!d) { set_error("b200cc_gemm_tf32x3: null descriptor"); return 1; }
  if (d->struct_size != (int)sizeof(b200cc_gemm3_desc)) {
    set_error("b200cc_gemm_tf32x3: descriptor is %d bytes, this library's b200cc_gemm3_desc is %d (stale binding? see include/b200cc.h)",
              d->struct_size, (int)sizeof(b200cc_gemm3_desc));
    return 1;
  }
  if (d->M <= 0 || d->N <= 0 || d->batch <= 0) return 0;
  if (d->K <= 0) { set_error("b200cc_gemm_tf32x3: K must be positive"); return 1; }
  if (!d->Ahi || !d->Alo || !d->Bhi || !d->Blo || !d->C) { set_error("b200cc_gemm_tf32x3: null operand"); return 1; }
  if (!mx::al16(d->Ahi) || !mx::al16(d->Alo) || !mx::al16(d->Bhi) || !mx::al16(d->Blo) || (d->lda & 3) || (d->ldb & 3) ||
      (d->strideA & 3) || (d->strideB & 3) || d->lda < d->K || d->ldb < d->K) {
    set_error("b200cc_gemm_tf32x3: planes must be 16-byte aligned with pitches / strides that are multiples of 4 floats");
    return 1;
  }
  if (d->K2 > 0 && (!d->A2hi || !d->A2lo || !d->B2hi || !d->B2lo || !mx::al16(d->A2hi) || !mx::al16(d->A2lo) ||
                    !mx::al16(d->B2hi) || !mx::al16(d->B2lo) || (d->lda2 & 3) || (d->ldb2 & 3) || (d->strideA2 & 3) ||
                    (d->strideB2 & 3) || d->lda2 < d->K2 || d->ldb2 < d->K2)) {
    set_error("b200cc_gemm_tf32x3: bad second-segment planes");
    return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (d->config) {
    case 0:
    case 5: return mx::launch<128, 32, 3, true>(d, st);
    case 6: return mx::launch<128, 16, 6, true>(d, st);
    case 7: return mx::launch_r2<32, 4>(d, st);
    case 1: return mx::launch<256, 32, 2, false>(d, st);
    case 2: return mx::launch<128, 32, 3, false>(d, st);
    case 3: return mx::launch<256, 16, 4, false>(d, st);
    case 4: return mx::launch<128, 16, 6, false>(d, st);
    default: set_error("b200cc_gemm_tf32x3: unknown config %d", d->config); return 1;
  }
}
