/*
 * b200cc.h -- C ABI of libb200cc.so: hand-written sm_100a kernels for the closed-shell
 * RHF-CCSD amplitude iteration and the (T) correction of CrawfordGroup/pycc.
 *
 * The reference (pure Python) has no FFI; its single contraction choke point is
 * ContractionBackend.__call__ (pycc/device.py:64-86, opt_einsum -> torch.tensordot -> cuBLAS),
 * plus ATen elementwise ops and the Python loops of cctriples.py.  Each entry point below names
 * the reference lines it replaces.  Conventions:
 *   - all pointers are DEVICE pointers to float64 unless stated otherwise; the caller (torch) owns
 *     every buffer, the library owns nothing but a small per-device scratch for reductions;
 *   - sizes/strides are in ELEMENTS (doubles), `long long`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     stream-ordered, no call synchronises unless stated;
 *   - return value 0 = OK, non-zero = error, text via b200cc_last_error() (thread-local).
 * There is NO CPU implementation behind any of these: without a CUDA device every compute call
 * fails with a non-zero status.
 */
#ifndef B200CC_H
#define B200CC_H

#ifdef __cplusplus
extern "C" {
#endif

typedef long long b200cc_i64;

int b200cc_version(void);          /* 200 = this header (round 2: struct_size fields, pair-packed ladder entries) */
const char* b200cc_last_error(void);
/* number of kernels launched by this library in this process (for bench.py's gpu_launches) */
b200cc_i64 b200cc_launch_count(void);
int b200cc_device_info(int* sm_count, int* cc_major, int* cc_minor, b200cc_i64* free_bytes, b200cc_i64* total_bytes);

/* ---- FP64 tensor-core (DMMA m8n8k4) GEMM -----------------------------------------------------
 * Replaces every two-operand contraction that opt_einsum lowers to tensordot/cuBLAS DGEMM
 * (pycc/device.py:84), incl. the particle-particle ladder (ccwfn.py:931), the ring terms
 * (ccwfn.py:644-645,683,715,933-938), Wmnij (603,930) and the t3 build (cctriples.py:50-62).
 *
 *   C[b](m,n) = alpha * ( sum_k opA1[b](m,k) opB1[b](n,k)  [+ sum_k opA2[b](m,k) opB2[b](n,k)] ) + beta * C[b](m,n)
 *
 * transA = 0: A is "K-major", element (m,k) at A[m*lda + k];  transA = 1: "M-major", A[k*lda + m].
 * transB = 0: B is "K-major", element (n,k) at B[n*ldb + k];  transB = 1: "N-major", B[k*ldb + n].
 * C is row-major (m,n) at C[m*ldc + n].  The optional second K-segment (K2 > 0) accumulates a second
 * product into the same tile before the epilogue (used by (T): particle + hole term in one pass).
 * Batching: either constant strides (strideA1.., `table` NULL) or a device table of
 * 5 x batch absolute addresses {A1,B1,A2,B2,C} (int64) per batch entry.
 * ksplit > 1 splits the K loop over gridDim.z and reduces through `workspace`
 * (>= ksplit*batch*M*N doubles); deterministic (no atomics).                                    */
typedef struct {
  int struct_size;                 /* = sizeof(b200cc_gemm_desc): a binding built against another layout of this struct
                                      is refused with an error instead of being mis-read */
  int M, N;
  int transA, transB;
  int K1, K2;                      /* K2 = 0: single segment */
  const double *A1, *B1, *A2, *B2;
  b200cc_i64 lda1, ldb1, lda2, ldb2;
  b200cc_i64 strideA1, strideB1, strideA2, strideB2;
  double* C;
  b200cc_i64 ldc, strideC;
  double alpha, beta;
  int batch;
  const b200cc_i64* table;         /* device, [batch][5] or NULL */
  int table_align16;               /* table mode: caller vouches every address is 16-byte aligned */
  int ksplit;
  double* workspace;
  int out_cube_nv;                 /* > 0: C is written 8x8x8-cube-blocked with row = x*nv + y, col = z (see b200cc_t_q_size);
                                      TMA kernels only, beta must be 0 */
  const int* bcoords;              /* device, [batch][4] or NULL: per-batch INDEX of {A1,B1,A2,B2} along each operand's own
                                      batch stride (counts nbA1..nbB2); C stays strided.  K-major TMA kernels only ((T)) */
  int nbA1, nbB1, nbA2, nbB2;
  int config;                      /* kernel: 0 auto (tile by a cost model, TMA when eligible); 2 = plain 128x128 multistage kernel (8 warps,
                                      __syncthreads pipeline; kept as a cross-check); 4 = 128x128 warp-specialised
                                      (8 DMMA warps + 4 cp.async producer warps, mbarrier pipeline); 5 = 80x128 ditto;
                                      6 / 7 = as 4 / 5 with cp.async.bulk.tensor (TMA) operand staging, K-major aligned operands only */
  /* optional THIRD and FOURTH K segment accumulated into the same tile (K3 = 0: none; K4 needs K3, K3 needs K2) -- TMA
     kernels only.  Used by (T) to sum two of the six t3 products into ONE output array (W needs them at transposed row
     pairs: the second product reads a constant copy of its operand with the pair transposed), which halves the t3
     traffic through HBM.  With bcoords set and K3 > 0 a batch entry carries 8 slab indices {A1,B1,A2,B2,A3,B3,A4,B4}. */
  int K3, K4;
  const double *A3, *B3, *A4, *B4;
  b200cc_i64 lda3, ldb3, lda4, ldb4;
  b200cc_i64 strideA3, strideB3, strideA4, strideB4;
  int nbA3, nbB3, nbA4, nbB4;
} b200cc_gemm_desc;

int b200cc_dgemm(const b200cc_gemm_desc* d, void* stream);

/* ---- mixed precision: split-TF32 GEMM on the tcgen05 tensor cores, FP64 accumulation -------------
 * precision='MP' replacement for the large K-major x K-major contractions of b200cc_dgemm (same reference
 * lines: ccwfn.py:931 ladder, 644-645/683/715/933-935 ring terms, 603 Wmnij).  The reference has only
 * whole-calculation float32 ('SP', device.py:147-151); this mode keeps every tensor FP64 and evaluates
 *
 *   C[b](m,n) = alpha * sum_k A[b](m,k) B[b](n,k) + beta * C[b](m,n)
 *
 * as Ahi.Bhi + Ahi.Blo + Alo.Bhi with TF32 operands (hi + lo ~= the FP64 value to 2^-22), FP32 accumulation in
 * TMEM over at most `kchunk` summation indices at a time and FP64 accumulation of the chunks in C.
 *
 * b200cc_split_tf32: src (rows x K, pitch ld doubles, `batch` copies `stride` doubles apart) ->
 * hi/lo FP32 planes [batch][rows][ldp], ldp % 4 == 0, ldp >= K, columns K..ldp zeroed, 16-byte aligned.
 * b200cc_gemm_tf32x3: planes are K-major: A element (m,k) at Ahi[b*strideA + m*lda + k] (floats; lda, ldb,
 * strideA, strideB multiples of 4; stride 0 = one operand shared by the whole batch); C row-major FP64.        */
typedef struct {
  int struct_size;                 /* = sizeof(b200cc_gemm3_desc), checked like b200cc_gemm_desc.struct_size */
  int M, N, K;
  const float *Ahi, *Alo, *Bhi, *Blo;
  b200cc_i64 lda, ldb, strideA, strideB;
  double* C;
  b200cc_i64 ldc, strideC;
  double alpha, beta;
  int batch;
  int kchunk;                      /* summation indices per FP32 (TMEM) accumulation chunk; 0 = 256 (configs 5/6) / 2048 (1..4) */
  int config;                      /* 0 auto = 5; 5 / 6 = 128x128 tile, FP64 running sums in registers, separate TMEM accumulators
                                      for the hi.hi and the cross products, 32- / 16-wide k-blocks (3 / 6 stages);
                                      1..4 = one shared accumulator drained into C in global memory:
                                      1 = 128x256 tile, 32-wide k-blocks, 2 stages; 2 = 128x128, 32, 3 stages;
                                      3 = 128x256, 16-wide k-blocks (64-byte swizzle), 4 stages; 4 = 128x128, 16, 6 stages */
  /* optional second K segment accumulated into the same tile (the (T) hole term), configs 0/5/6 only; K2 = 0: none */
  int K2;
  const float *A2hi, *A2lo, *B2hi, *B2lo;
  b200cc_i64 lda2, ldb2, strideA2, strideB2;
  const int* bcoords;              /* device, [batch][4] or NULL: per-batch slab INDEX of {A1,B1,A2,B2} along each operand's own
                                      stride (slab counts nbA1..nbB2); C stays strided (as b200cc_gemm_desc.bcoords) */
  int nbA1, nbB1, nbA2, nbB2;
  int lockstep;                    /* configs 5/6: leader/follower tile schedule -- a CTA requests an operand tile only after the
                                      CTA that leads its row / column of the tile block has received it, so every tile is read from
                                      DRAM once and served from L2 afterwards.  > 0: on, followers run this many k-blocks behind;
                                      <= 0: off (default: measured to cut DRAM traffic but not run time, see mixed.cu) */
} b200cc_gemm3_desc;

int b200cc_split_tf32(const double* src, b200cc_i64 ld, b200cc_i64 stride, int rows, int K, int batch,
                      float* hi, float* lo, b200cc_i64 ldp, void* stream);
int b200cc_gemm_tf32x3(const b200cc_gemm3_desc* d, void* stream);
/* b200cc_merge_tf32: the inverse of b200cc_split_tf32 for one batch entry -- dst (rows x K doubles, pitch ld) = hi + lo.
 * Used where precision='MP' has released the FP64 <ab|ef> block and an FP64-only consumer needs rows of it back
 * (t_if <ab|ef> inside HBAR's Hvvvo, cchbar.py:632-688; the CC2 / CC3 t1-dressed <ab|ef> terms, ccwfn.py:880, 1104);
 * the result carries the 2^-22 relative accuracy of the planes, i.e. that of the mode.                          */
int b200cc_merge_tf32(const float* hi, const float* lo, b200cc_i64 ldp, b200cc_i64 rows, int K, double* dst,
                      b200cc_i64 ld, void* stream);

/* ---- the particle-particle ladder in symmetric / antisymmetric pair form ------------------------------------
 * Replaces 'ijef,abef->ijab' (ccwfn.py:931; the Lambda ladder cclambda.py:468 likewise) as written -- a GEMM with
 * N = K = v^2 over the v^4 block <ab|ef> -- by the split the reference itself uses for W_abei (ccwfn.py:1054-1120).
 * With pair(x,y) = x(x+1)/2 + y, P = pair(a,b) (a >= b), Q = pair(e,f) (e >= f), nq = v(v+1)/2:
 *   V+[P,Q] = <ab|ef> + <ab|fe> (e > f), <ab|ee> (e = f);      V-[P,Q] = <ab|ef> - <ab|fe>   (zero for a = b, e = f)
 *   T+[m,Q] = (tau[m,e,f] + tau[m,f,e]) / 2 (tau[m,e,e]);      T-[m,Q] = (tau[m,e,f] - tau[m,f,e]) / 2
 *   S = T+ V+^T,  A = T- V-^T  (b200cc_dgemm / b200cc_gemm_tf32x3, batch of 2, N = K = nq),
 *   sum_ef tau[m,e,f] <ab|ef> = S[m,P] + A[m,P],   sum_ef tau[m,e,f] <ba|ef> = S[m,P] - A[m,P].
 * `tri` != 0: tau has the pair symmetry tau[i,j,e,f] = tau[j,i,f,e] (the amplitudes solve_cc iterates), rows are
 * m = pair(i,j), i >= j only, and the (j,i) results follow from (i,j): o^2 v^4 / 2 executed flop instead of 2 o^2 v^4.
 * `tri` == 0: m = i*no + j over all (i,j), any tau (o^2 v^4 executed flop).
 * All matrices are row-major with pitch ldq (>= nq; the columns nq..ldq of packed rows are written as zeros).
 *
 * b200cc_pack_pairs: rows of <ab|ef> for a in [a0,a1) -> V+ / V- rows pair(a,b) - pair(a0,0), b <= a.  The source is
 *   addressed by element strides, slab(a,b)[e,f] = src[(a-a0)*sa + b*sb + e*se + f*sf] (so the output of the
 *   integral-generating GEMM is packed in whatever index order it was produced, no permuted copy).  Once per Hamiltonian.
 * b200cc_unpack_pairs: the inverse for `npairs` consecutive rows: dst[p][e][f] = <ab|ef> of the p-th given pair (FP64,
 *   v^2 per pair) -- for the few other consumers of <ab|ef> (t_if<ab|ef> in HBAR / CC2 / CC3, H.ERI[v,v,v,v]).
 * b200cc_pack_tau: tau (no,no,nv,nv) -> T+ / T- (every iteration; also used for lambda_2 by the Lambda ladder).
 * b200cc_ladder_unpack: r2[i,j,a,b] += alpha (S+A), r2[i,j,b,a] += alpha (S-A) [and the (j,i) images when tri] for the
 *   columns pair(a,b) - pair(a0,0), a in [a0,a1) of S / A (pitch lds): each r2 element is updated by exactly one thread. */
b200cc_i64 b200cc_pair_count(int n);          /* n(n+1)/2 */
int b200cc_pack_pairs(const double* src, b200cc_i64 sa, b200cc_i64 sb, b200cc_i64 se, b200cc_i64 sf, int nv,
                      int a0, int a1, double* vp, double* vm, b200cc_i64 ldq, void* stream);
int b200cc_unpack_pairs(const double* vp, const double* vm, b200cc_i64 ldq, int nv, b200cc_i64 npairs,
                        double* dst, void* stream);
int b200cc_pack_tau(const double* tau, int no, int nv, int tri, double* tp, double* tm, b200cc_i64 ldq, void* stream);
int b200cc_ladder_unpack(const double* S, const double* A, b200cc_i64 lds, int no, int nv, int tri, int a0, int a1,
                         double alpha, double* r2, void* stream);
/* The same pair form for Z_mbij = <mb|ef> tau_ijef (ccwfn.py:715) with pair-symmetric tau:
 *   X+-[(m,b),Q] = <mb|ef> +- <mb|fe>  (b200cc_pack_rows: `nrows` (v,v) slabs of a constant block, once per Hamiltonian),
 *   S = T+ X+^T, A = T- X-^T over the rows pair(i,j) (the T+- of the ladder),  Z[i,j] = S + A,  Z[j,i] = S - A
 * (b200cc_pair_rows_unpack: out[(i*no+j)*ldo + c] = S[p,c] + A[p,c], out[(j*no+i)*ldo + c] = S[p,c] - A[p,c]):
 * o^3v^3 executed flop instead of 2 o^3v^3.                                                                       */
int b200cc_pack_rows(const double* src, b200cc_i64 nrows, int nv, double* xp, double* xm, b200cc_i64 ldq, void* stream);
int b200cc_pair_rows_unpack(const double* S, const double* A, b200cc_i64 lds, int no, b200cc_i64 ncols, double* out,
                            b200cc_i64 ldo, void* stream);

/* The two "ring layouts" of t2 that the o^3v^3 products of the residual read (ccwfn.py:933-935), in ONE pass over t2:
 *   u[i,a,m,e] = 2 t2[i,m,a,e] - t2[i,m,e,a],   tb[i,a,m,e] = t2[i,m,e,a]      (both (no,nv,no,nv), contiguous)      */
int b200cc_ring_layouts(const double* t2, int no, int nv, double* u, double* tb, void* stream);

/* ---- tensor permutation / strided axpby -------------------------------------------------------
 * out[sum_d i_d*so[d]] = alpha * in[sum_d i_d*si[d]] + beta * out[...]  for i_d < shape[d], rank <= 6.
 * Replaces tensordot's permute+contiguous copies and swapaxes/clone/+ ATen passes
 * (e.g. ccwfn.py:757,759,922,933-934).  beta == 0 never reads `out`.                            */
int b200cc_permute(int rank, const b200cc_i64* shape, const b200cc_i64* si, const b200cc_i64* so,
                   double alpha, const double* in, double beta, double* out, void* stream);

/* z = a*x + b*y  (n elements; z may alias x or y).  helper_diis error vectors (utils.py:291-293). */
int b200cc_axpbyz(b200cc_i64 n, double a, const double* x, double b, const double* y, double* z, void* stream);

/* tau = f1*t2 + f2*t1(x)t1   (build_tau, ccwfn.py:455) */
int b200cc_build_tau(int no, int nv, double f1, double f2, const double* t1, const double* t2,
                     double* tau, void* stream);

/* out[i,j,a,b] = in[i,j,a,b] / (eo[i]+eo[j]-ev[a]-ev[b])  (ccwfn.py:195-198,211); in-place allowed.
 * t1 form: out[i,a] = in[i,a] / (eo[i]-ev[a]).                                                   */
int b200cc_div_d2(int no, int nv, const double* eo, const double* ev, const double* in, double* out, void* stream);
int b200cc_div_d1(int no, int nv, const double* eo, const double* ev, const double* in, double* out, void* stream);

/* Fused Jacobi update of solve_cc (ccwfn.py:281-284) with the r2 symmetrisation of r_T2 (ccwfn.py:790):
 *   r2 = half + half^T(ij<->ji, ab<->ba);  t1 += r1/Dia;  t2 += r2/Dijab;
 *   sumsq[0] = sum (r1/Dia)^2 + sum (r2/Dijab)^2          (device scalar; rms = sqrt on host)
 * `r2_half` is overwritten with the symmetrised r2 when write_r2 != 0.  `symmetrize` = 0 treats
 * r2_half as already symmetric.  scratch: >= 4096 doubles.                                       */
int b200cc_update_amps(int no, int nv, const double* eo, const double* ev, const double* r1,
                       double* r2_half, int symmetrize, int write_r2, double* t1, double* t2,
                       double* sumsq, double* scratch, void* stream);

/* The same update for the rows i in [i0,i1) of t2 only -- one process per GPU: after the all-reduce of the half
 * residual each rank updates its rows, then the rows are all-gathered (parallel.py).  t1 (replicated, tiny) is updated
 * in full by every caller when r1 != NULL.  sumsq2[0] = sum (r2/D)^2 over the rows, sumsq2[1] = sum (r1/D)^2.
 * r2_half is not modified.  scratch: >= 4096 doubles.                                                            */
int b200cc_update_amps_rows(int no, int nv, int i0, int i1, const double* eo, const double* ev, const double* r1,
                            const double* r2_half, double* t1, double* t2, double* sumsq2, double* scratch,
                            void* stream);

/* r2 = half + half^T in place (r_T2, ccwfn.py:790) */
int b200cc_symmetrize_r2(int no, int nv, double* r2, void* stream);

/* E = 2 sum f_ia t_ia + sum (t2 + t1 t1)_ijab L_ijab   (cc_energy, ccwfn.py:1160-1161).
 * fov: (no,nv) view with leading dimension ldf.  scratch: >= 4096 doubles.                       */
int b200cc_cc_energy(int no, int nv, const double* fov, b200cc_i64 ldf, const double* t1, const double* t2,
                     const double* Loovv, double* e_out, double* scratch, void* stream);

/* The rows i in [i0,i1) of the doubles part (+ the singles part when with_singles != 0): partial energies of the
 * ranks are summed by the caller.                                                                              */
int b200cc_cc_energy_rows(int no, int nv, int i0, int i1, int with_singles, const double* fov, b200cc_i64 ldf,
                          const double* t1, const double* t2, const double* Loovv, double* e_out, double* scratch,
                          void* stream);

/* out[q] = sum_i x[i]*y_q[i], q < m <= 16; `ys` is a HOST array of m device pointers.
 * One pass over x; deterministic two-stage reduction.  (helper_diis B matrix, utils.py:330-339;
 * rms contractions ccwfn.py:283.)  scratch: >= 16*1024 doubles.                                  */
int b200cc_multi_dot(b200cc_i64 n, const double* x, int m, const double* const* ys, double* out,
                     double* scratch, void* stream);

/* out = sum_q c[q]*xs[q], q < m <= 16; `xs` HOST array of device pointers, `c` HOST coefficients.
 * (helper_diis extrapolation, utils.py:353-355).  out may not alias any xs[q].                   */
int b200cc_multi_axpy(b200cc_i64 n, int m, const double* c, const double* const* xs, double* out, void* stream);

/* ---- (T): Lee-Rendell energy of a batch of (i>=j>=k) triples ---------------------------------
 * Replaces the per-triple epilogue of t_tjl (cctriples.py:208-237: t3d_ijk, the 1/(1+delta)
 * scaling, X3/Y3/Z3, denominators, the a>=b>=c sum).  Input: for each of `ntrip` triples the six
 * GEMM outputs Q1..Q6 (each (nv*nv) x nv) produced by b200cc_dgemm:
 *   W[a,b,c] = Q1[a,b,c]+Q2[a,c,b]+Q3[c,a,b]+Q4[c,b,a]+Q5[b,c,a]+Q6[b,a,c]
 * Q: device, [ntrip][6][nv^3].  ijk: device int32 [ntrip][3].  fov: (no,nv) view, leading dim ldf.
 * q_blocked is a set of layout flags: bit 0 = the arrays are cube-blocked (b200cc_t_q_size), bit 1 = PAIRED: only three
 * arrays per triple, [ntrip][3][nv^3], in which the GEMM (K3/K4 segments of b200cc_gemm_desc) has already summed
 *   R1[x,y,z] = Q1[x,y,z] + Q6[y,x,z],   R2[x,y,z] = Q2[x,y,z] + Q3[y,x,z],   R3[x,y,z] = Q4[x,y,z] + Q5[y,x,z],
 *   W[a,b,c] = R1[a,b,c] + R2[a,c,b] + R3[c,b,a]     (half the t3 bytes through HBM; same flags for b200cc_t3_assemble).
 * et_out[0] (+)= sum over the batch (accumulate != 0 adds to the existing value).
 * scratch >= number of CTAs doubles (b200cc_t_energy_scratch).                                   */
b200cc_i64 b200cc_t_energy_scratch(int nv, int ntrip);
/* doubles per Q array: nv^3 (q_blocked = 0), or ceil(nv/8)^3 * 512 when the GEMM wrote it as contiguous 8x8x8 cubes
 * (b200cc_gemm_desc.out_cube_nv): element (x,y,z) at cube (x/8,y/8,z/8), offset (x%8)*64 + (y%8)*8 + z%8. */
b200cc_i64 b200cc_t_q_size(int nv, int blocked);
int b200cc_t_energy_batch(int no, int nv, int ntrip, const int* ijk, const double* Q, int q_blocked,
                          const double* t1, const double* t2, const double* oovv,
                          const double* fov, b200cc_i64 ldf, const double* eo, const double* ev,
                          double* et_out, int accumulate, double* scratch, void* stream);

/* The connected (w3_out, = t3c_ijk) and disconnected (v3_out, = t3d_ijk) t3 numerators of ONE triple
 * as (nv,nv,nv) arrays, w3 assembled from Q1..Q6 -- the per-triple parity hook for
 * cctriples.py:27-72 and 108-147.  with_denom != 0 divides both by D_ijkabc.  v3_out may be NULL. */
int b200cc_t3_assemble(int no, int nv, int i, int j, int k, const double* Q, int q_blocked,
                       const double* t1, const double* t2, const double* oovv,
                       const double* fov, b200cc_i64 ldf, const double* eo, const double* ev,
                       int with_denom, double* w3_out, double* v3_out, void* stream);

/* ---- (T) densities and Lambda sources (cctriples.py:1063-1157, t3_density) --------------
 * Step 1: connected t3 WITH denominators for a batch of triples, from the six GEMM outputs per triple:
 *   m3_out[t][a,b,c] = (Q1[a,b,c]+Q2[a,c,b]+Q3[c,a,b]+Q4[c,b,a]+Q5[b,c,a]+Q6[b,a,c]) / D_ijkabc   (= t3c_ijk,
 *   cctriples.py:50-70).  Q: [ntrip][6][nv^3] (plain layout), ijk: device int32 [ntrip][3].                  */
int b200cc_t3_connected_batch(int no, int nv, int ntrip, const int* ijk, const double* Q, const double* eo,
                              const double* ev, double* m3_out, void* stream);

/* Step 2: for fixed (i,j) and k = k0 .. k0+nk-1, everything of the loop body of t3_density (cctriples.py:1121-1148)
 * that is not a GEMM.  With M3 from step 1, N3 = t3d_ijk/D formed on the fly, X3 = sym(M3), Y3 = sym(N3):
 *   GEMM operands (written, not accumulated):
 *     W2 = 2 X3 + Y3  ->  W2ab[(a,b)][kk][c]  and  W2n[a][kk][(b,c)]        (kk = k - k0)
 *     P  = 2 M3 - M3[a,c,b] - M3[c,b,a]  ->  Pab, Pn  (same two layouts)
 *   accumulated in place:
 *     Gij[a,b] += 4 t1[k,c] Z3[a,b,c]            (Goovv[i,j], line 1141)
 *     Xij[a,b] += (M3 - M3[c,b,a]) f[k,c]        (X2[i,j], line 1128)
 *     dvv[a]   += 1/2 M3 (X3 + Y3)               (line 1133; doo[i] = - sum_a of the same, line 1134)
 *     Dov[a]   += (M3 - M3[c,b,a]) (4 t2[j,k,b,c] - 2 t2[j,k,c,b])          (Dov[i], line 1137)
 *     S1[a]    += 2 (M3 - M3[b,a,c]) (2<jk|bc> - <jk|cb>)                   (S1[i], line 1146)
 * t2s = 4 t2 - 2 t2.swapaxes(2,3) and oovvs = 4<ij|ab> - 2<ij|ba> are (no,no,nv,nv) arrays built once by the caller
 * (sym() of the disconnected t3 needs only these combinations, see triples.cu).
 * swap_ab != 0: M3 holds the run of the transposed pair, M3[kk][b,a,c] = t3c(i,j,k)[a,b,c] (one t3 build serves
 * (i,j) and (j,i)).  scratch: >= b200cc_t3_density_scratch(nv) doubles.  Deterministic (no atomics).            */
typedef struct {
  int no, nv, i, j, k0, nk, swap_ab;
  const double* M3;                    /* [nk][nv^3] */
  const double *t1, *t2s, *oovvs, *fov; /* fov: (no,nv) view, leading dimension ldf */
  b200cc_i64 ldf;
  const double *eo, *ev;
  double *W2ab, *W2n, *Pab, *Pn;       /* nk*nv^3 each */
  double *Gij, *Xij;                   /* (nv,nv) */
  double *dvv, *Dov, *S1;              /* nv each */
  double* scratch;
} b200cc_t3d_desc;
b200cc_i64 b200cc_t3_density_scratch(int nv);
int b200cc_t3_density_forms(const b200cc_t3d_desc* d, void* stream);

/* ---- (T), fused (a,b,c)-driven form: the t3 tile never goes through HBM (csrc/triples_abc.cu) -------------------
 * Replaces, for a list of virtual triples a >= b >= c: t3c_abc (cctriples.py:75-105), t3d_abc (149-173) and the
 * Lee-Rendell bracket of t_tjl (208-237) with the roles of (i,j,k) and (a,b,c) exchanged -- same E(T).  One persistent
 * CTA per (a,b,c): three FP64 DMMA GEMMs [o^2 pairs] x [o], K = 2v + 2o, operands streamed by TMA, accumulated into the
 * CTA's private o^3 tile, followed by the energy evaluation of that tile in the same kernel.
 *   et_out[0] (+)= sum over the listed (a,b,c) of their E(T) contributions (weights 2 - d_ab - d_ac - d_bc included).
 * Needs even o <= b200cc_t_abc_max_no(), even v <= 1023.  Constant operands (built once per amplitude set):
 *   G[l][x][y][e]   = <le|xy>  (= ovvv[l,e,x,y], Wvvvo[y,x,e,l])          o v^3
 *   t2x[x][y][l][m] = t2[l,m,x,y]                                         o^2 v^2
 *   Ox[z][p][q][m]  = -<mz|pq> (= -Wovoo[m,z,p,q])                        o^3 v
 *   oovvx[x][y][i][j] = <ij|xy>                                           o^2 v^2
 * abc / sorted: device int32, one entry per triple packed as  x0 | x1 << 10 | x2 << 20  with x0 >= x1 >= x2
 * (sorted: the occupied triples i >= j >= k, i = j = k left out -- they contribute exactly zero).
 * wtile: grid * o^3 doubles of scratch (after a call with nabc = 1, grid = 1 it holds W_abc[i][j][k], the connected
 * numerator -- the parity hook for t3c_abc); partial: grid doubles.
 * While the job runs the scratch tiles are pinned in the L2 by a persisting access window (so the t3 tile is not written
 * back to HBM); the call waits for the job on `stream` and then releases the window and the L2 set-aside, i.e. unlike
 * the other entry points it returns after its kernels have finished (B200CC_TABC_L2PERSIST=0: no window, no wait). */
typedef struct b200cc_t_abc_desc {
  int struct_size;          /* sizeof(b200cc_t_abc_desc), checked by the library */
  int no, nv;
  int nabc, nsorted;
  const int* abc;
  const int* sorted;
  const double* G;
  const double* t2;
  const double* t2x;
  const double* Ox;
  const double* oovvx;
  const double* t1;
  const double* fov;        /* F[o,v] block, row pitch ldf */
  b200cc_i64 ldf;
  const double* eo;
  const double* ev;
  double* wtile;
  double* partial;
  double* et_out;
  int accumulate;
  int grid;                 /* CTAs to launch (one per SM); the scratch arrays are sized by it */
  int fov_is_zero;          /* caller vouches that F[o,v] = 0 (canonical reference): its terms of t3d_abc are not evaluated */
} b200cc_t_abc_desc;
int b200cc_t_abc_max_no(void);
int b200cc_t_abc(const b200cc_t_abc_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif
