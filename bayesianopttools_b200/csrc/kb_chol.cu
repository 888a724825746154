// krig-b200 K2: population-batched blocked FP64 Cholesky on the FP64 tensor pipe.
//
// Replaces np.linalg.cholesky (likelihood_func.py:175) and the forward halves of
// the six triangular solves (likelihood_func.py:183-193) for every candidate of
// the hyper-parameter population at once.
//
// Design (B200-first, not LAPACK-shaped):
//  * Each candidate's matrix is the AUGMENTED lower-trapezoid [R ; (y F)^T ; I?]:
//    the right-hand sides ride along as extra tile ROWS, so the panel solve of
//    the factorisation also produces L^-1 y, L^-1 F (and L^-T in fit mode), and
//    the trailing Schur tile is -(L^-1[y F])^T (L^-1[y F]) = the GLS normal
//    equations.  No separate, latency-bound forward substitution exists.
//  * Left-looking 64x64 TILE tasks  C = A[i,k] - sum_{p<k} L[i,p] L[k,p]^T  (DMMA m8n8k4, the
//    accumulator stays in registers for the whole K loop) followed by one DMMA multiply with the
//    inverted diagonal block.  One persistent launch for the whole population; CTAs split in
//    two roles so that the serial chain never queues behind (or shares an FP64 pipe with) bulk work:
//      - CHAIN workers (the CTAs of a few SMs) walk the critical path of their candidates,
//        step k:  L[k,k-1] = C1 Dinv_{k-1}^T ;  C2 -= L[k,k-1] L[k,k-1]^T ;  factor + invert
//        C2 -> L[k,k], Dinv_k.  C1/C2 arrive pre-updated with all terms p < k-1 from a bulk
//        PRE task, so a step is three tile loads, two 64^3 DMMA products and the in-shared
//        factorisation (kb_diag.cuh) - no global round trip inside the chain.
//      - BULK workers pull tile tasks from a global queue ordered so that every dependency
//        is EARLIER in the queue; completion is published through per-row progress words
//        (st.release.gpu / ld.acquire.gpu), polled only when the K loop catches up with them.
//        The oldest unfinished task always has all inputs (chain steps run as soon as their PRE
//        task is done) -> no deadlock and no grid-wide barrier.
//  * Operand tiles stream L2 -> shared memory through a 3-stage cp.async ring; shared tiles
//    are padded (row stride == 4 mod 16 doubles) so every DMMA fragment load is conflict free.
//  * Wasted tensor work is skipped: only the lower 8x8 blocks of diagonal tiles are
//    accumulated, the multiply by the (triangular) inverse stops at the diagonal, and when the
//    right-hand sides occupy <= 16 rows (ordinary Kriging: y and 1) they run as THIN tasks
//    (16 x 64 outputs, one 8-column block per warp).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kb_common.cuh"
#include "kb_internal.h"
#include "kb_diag.cuh"

// K chunk (doubles) per pipeline stage and stage count, per task shape (compile-time, -D overridable for A/B builds).
// tools/ubench/kloop.cu runs the K loops alone on the real access stream (no dependencies, no epilogues; operands from
// HBM or L2 make no difference): 128-row tasks 32 wide x 2 stages reach 32.0 TFLOP/s of the 37.1 peak, 16 wide x 3 stages
// 34.2, with the copies compiled out 35.6 (profiles/r02_ubench_kloop.txt).  Inside the factorisation the two rings are
// equal within noise (n=1000: 1.118 vs 1.135 ms, n=2000: 6.63 vs 6.66 ms): the tasks are not K-loop-throughput bound
// (DESIGN 5c), so the wider chunk with fewer barriers stays.
#ifndef KB_KC128
#define KB_KC128 32     // 128-row tasks (MODE 4)
#define KB_NS128 2
#endif
#ifndef KB_KC64
#define KB_KC64 32      // 64-row, PRE and thin tasks (MODE 0, 1, 2)
#define KB_NS64 3
#endif

// One ring of 384 x 36 doubles: the pipeline stages of the K loop (rows padded by 4 doubles: row stride == 4 mod 16
// doubles for KC = 16 and 32, conflict-free fragment loads).  Epilogue tiles and the chain worker's three tiles alias it.
#define KB_RING_DOUBLES (384 * 36)
#ifndef KB_EARLY_CLAIM
#define KB_EARLY_CLAIM 0
#endif
#ifndef KB_CHOL_HELPERS_DEFAULT
#define KB_CHOL_HELPERS_DEFAULT 3    // bulk CTAs run chain steps at the start (1) and at the end (2) of the launch
#endif
#ifndef KB_CHOL_GROUPS_DEFAULT
#define KB_CHOL_GROUPS_DEFAULT 1     // candidate groups of the staggered queue (kb_chol_stagger)
#endif
static_assert(KB_NS128 * (128 + 64) * (KB_KC128 + 4) <= KB_RING_DOUBLES && KB_NS64 * (64 + 64) * (KB_KC64 + 4) <= KB_RING_DOUBLES,
              "pipeline stages must fit in the ring");
struct KbCholSmem {
  double ring[KB_RING_DOUBLES];
  int4 tdesc;
  int task;
  int ready;
  int misc[6];
  unsigned long long poll_ns;   // trace aid: time thread 0 spent polling progress words inside this task's K loop
  int knext[8];       // chain worker: next step of each home candidate, as far as this worker knows
  unsigned long long argbuf[48];   // copy of the kernel arguments for the out-of-line chain step (the kernel's own
                                   // parameter block stays in the constant bank for the inlined tile tasks)
};
static_assert(sizeof(double) * KB_RING_DOUBLES >= sizeof(double) * (3 * KB_TILE * KB_LDC + KB_DIAG_SCRATCH),
              "three padded tiles and the exchange scratch must fit in the pipeline ring");

__device__ __forceinline__ unsigned long long kb_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// 64 x 64 tile (row stride ld in global) -> padded shared tile, 16-byte cp.async per thread.
__device__ __forceinline__ void kb_load_tile_async(double (*dst)[KB_LDC], const double* src, int ld) {
#pragma unroll
  for (int j = 0; j < (KB_TILE * KB_TILE / 2) / KB_THREADS; ++j) {
    const int idx = threadIdx.x + j * KB_THREADS;
    const int row = idx >> 5, seg = (idx & 31) * 2;
    kb_cp_async16(&dst[row][seg], src + (size_t)row * ld + seg);
  }
}

// out = C * Dinv^T for a 64-row tile, Dinv lower triangular: column block cb only needs
// k < 8 (cb + 1).  Warp (wr, wc) owns rows wr*16.. and the interleaved column blocks
// cb = 2 ni + wc so both column halves carry the same work.
template <int MI>
__device__ __forceinline__ void kb_tri_multiply(const double (*Cs)[KB_LDC], const double (*Ds)[KB_LDC],
                                                double (&acc)[4][4][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, tq = lane & 3, wr = warp >> 1, wc = warp & 1;
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll
  for (int ks = 0; ks < KB_TILE / 4; ++ks) {
    double af[MI];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) af[mi] = Cs[wr * (8 * MI) + mi * 8 + gq][ks * 4 + tq];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      if (ks < 2 * (2 * ni + wc + 1)) {
        const double bf = Ds[(2 * ni + wc) * 8 + gq][ks * 4 + tq];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) kb_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf);
      }
    }
  }
}

// The same product with the A operand read from global memory (through L2): X0 has just been stored to the
// output tile by this CTA, and shared memory has no room for a fourth tile.  Rare path (see kb_refined_solve).
template <int MI>
__device__ __forceinline__ void kb_tri_multiply_g(const double* __restrict__ Xg, int ld, const double (*Ds)[KB_LDC],
                                                  double (&acc)[4][4][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, tq = lane & 3, wr = warp >> 1, wc = warp & 1;
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll 4
  for (int ks = 0; ks < KB_TILE / 4; ++ks) {
    double af[MI];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) af[mi] = __ldcg(&Xg[(size_t)(wr * (8 * MI) + mi * 8 + gq) * ld + ks * 4 + tq]);
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      if (ks < 2 * (2 * ni + wc + 1)) {
        const double bf = Ds[(2 * ni + wc) * 8 + gq][ks * 4 + tq];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) kb_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf);
      }
    }
  }
}

// X = C L^-T with one step of residual correction,
//     X0 = C W^T,   X1 = X0 + (C - X0 L^T) W^T          (W = the explicit inverse of the diagonal tile L),
// for diagonal tiles whose pivots span orders of magnitude (flag planted in the stored inverse by the chain).  The plain product with the
// explicit inverse leaves a residual of the order eps cond(L) |C|; on such tiles (the FIRST tile column of a
// smooth-kernel design with a tiny nugget: pivots from 1 down to the nugget) that put the likelihood 3-30x
// further from an 80-bit evaluation than LAPACK's triangular solve; with the correction it is as close as LAPACK
// (tools/emulate_blocking.py, profiles/r02_emulate_blocking.txt).  Well-conditioned tiles -- every tile of the
// reference's default nugget on space-filling designs -- take the single product.
// On entry Cs holds C and Bs holds W; Wg / Lg are the global copies of W and L (L with row stride ldl); X0 goes
// through the output tile Xg (row stride ld).  On exit Xg (and the shared tile Xs, when given; it may alias Cs)
// holds X1; Cs and Bs are clobbered.  All 256 threads; out of line so that the common path keeps its registers.
template <int MI>
__device__ __noinline__ void kb_refined_solve(double (*Cs)[KB_LDC], double (*Bs)[KB_LDC], const double* Wg,
                                              const double* Lg, int ldl, double* Xg, int ld, double (*Xs)[KB_LDC]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, tq = lane & 3, wr = warp >> 1, wc = warp & 1;
  double acc[4][4][2];
  kb_tri_multiply<MI>(Cs, Bs, acc);
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int r = wr * (8 * MI) + mi * 8 + gq, c = (2 * ni + wc) * 8 + 2 * tq;
      *reinterpret_cast<double2*>(&Xg[(size_t)r * ld + c]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    }
  __syncthreads();
  kb_load_tile_async(Bs, Lg, ldl);
  kb_cp_async_commit();
  kb_cp_async_wait<0>();
  __syncthreads();
  kb_tri_multiply_g<MI>(Xg, ld, Bs, acc);
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int r = wr * (8 * MI) + mi * 8 + gq, c = (2 * ni + wc) * 8 + 2 * tq;
      double2 v = *reinterpret_cast<double2*>(&Cs[r][c]);
      v.x -= acc[mi][ni][0];
      v.y -= acc[mi][ni][1];
      *reinterpret_cast<double2*>(&Cs[r][c]) = v;
    }
  __syncthreads();
  kb_load_tile_async(Bs, Wg, KB_TILE);
  kb_cp_async_commit();
  kb_cp_async_wait<0>();
  __syncthreads();
  kb_tri_multiply<MI>(Cs, Bs, acc);
  if (Xs) __syncthreads();                      // every warp is done reading Cs (== Xs) as an operand
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int r = wr * (8 * MI) + mi * 8 + gq, c = (2 * ni + wc) * 8 + 2 * tq;
      const double2 x0 = __ldcg(reinterpret_cast<const double2*>(&Xg[(size_t)r * ld + c]));
      const double2 x1 = make_double2(x0.x + acc[mi][ni][0], x0.y + acc[mi][ni][1]);
      *reinterpret_cast<double2*>(&Xg[(size_t)r * ld + c]) = x1;
      if (Xs) *reinterpret_cast<double2*>(&Xs[r][c]) = x1;
    }
}

// The same for a 16-row strip of right-hand sides (THIN tasks): warp w owns column block w.
__device__ __noinline__ void kb_refined_solve_thin(double (*Cs)[KB_LDC], double (*Bs)[KB_LDC], const double* Wg,
                                                   const double* Lg, int ldl, double* Xg, int ld) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int kmax = 2 * (warp + 1);
  double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  for (int ks = 0; ks < kmax; ++ks) {
    const double bf = Bs[warp * 8 + gq][ks * 4 + tq];
    kb_dmma(o[0][0], o[0][1], Cs[gq][ks * 4 + tq], bf);
    kb_dmma(o[1][0], o[1][1], Cs[8 + gq][ks * 4 + tq], bf);
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
    *reinterpret_cast<double2*>(&Xg[(size_t)(mi * 8 + gq) * ld + warp * 8 + 2 * tq]) = make_double2(o[mi][0], o[mi][1]);
  __syncthreads();
  kb_load_tile_async(Bs, Lg, ldl);
  kb_cp_async_commit();
  kb_cp_async_wait<0>();
  __syncthreads();
  double p[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  for (int ks = 0; ks < kmax; ++ks) {
    const double bf = Bs[warp * 8 + gq][ks * 4 + tq];
    kb_dmma(p[0][0], p[0][1], __ldcg(&Xg[(size_t)gq * ld + ks * 4 + tq]), bf);
    kb_dmma(p[1][0], p[1][1], __ldcg(&Xg[(size_t)(8 + gq) * ld + ks * 4 + tq]), bf);
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    double2 v = *reinterpret_cast<double2*>(&Cs[mi * 8 + gq][warp * 8 + 2 * tq]);
    v.x -= p[mi][0];
    v.y -= p[mi][1];
    *reinterpret_cast<double2*>(&Cs[mi * 8 + gq][warp * 8 + 2 * tq]) = v;
  }
  __syncthreads();
  kb_load_tile_async(Bs, Wg, KB_TILE);
  kb_cp_async_commit();
  kb_cp_async_wait<0>();
  __syncthreads();
  double dl[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  for (int ks = 0; ks < kmax; ++ks) {
    const double bf = Bs[warp * 8 + gq][ks * 4 + tq];
    kb_dmma(dl[0][0], dl[0][1], Cs[gq][ks * 4 + tq], bf);
    kb_dmma(dl[1][0], dl[1][1], Cs[8 + gq][ks * 4 + tq], bf);
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
    *reinterpret_cast<double2*>(&Xg[(size_t)(mi * 8 + gq) * ld + warp * 8 + 2 * tq]) =
        make_double2(o[mi][0] + dl[mi][0], o[mi][1] + dl[mi][1]);
}

// ... and for the first 8 * nmi rows of a tile (short tasks, nmi <= 7): warp w owns column block w.
__device__ __noinline__ void kb_refined_solve_short(double (*Cs)[KB_LDC], double (*Bs)[KB_LDC], const double* Wg,
                                                    const double* Lg, int ldl, double* Xg, int ld, int nmi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int kmax = 2 * (warp + 1);
  double o[7][2];
#pragma unroll
  for (int mi = 0; mi < 7; ++mi) o[mi][0] = o[mi][1] = 0.0;
  for (int ks = 0; ks < kmax; ++ks) {
    const double bf = Bs[warp * 8 + gq][ks * 4 + tq];
#pragma unroll
    for (int mi = 0; mi < 7; ++mi)
      if (mi < nmi) kb_dmma(o[mi][0], o[mi][1], Cs[mi * 8 + gq][ks * 4 + tq], bf);
  }
#pragma unroll
  for (int mi = 0; mi < 7; ++mi)
    if (mi < nmi)
      *reinterpret_cast<double2*>(&Xg[(size_t)(mi * 8 + gq) * ld + warp * 8 + 2 * tq]) = make_double2(o[mi][0], o[mi][1]);
  __syncthreads();
  kb_load_tile_async(Bs, Lg, ldl);
  kb_cp_async_commit();
  kb_cp_async_wait<0>();
  __syncthreads();
  {
    double p[7][2];
#pragma unroll
    for (int mi = 0; mi < 7; ++mi) p[mi][0] = p[mi][1] = 0.0;
    for (int ks = 0; ks < kmax; ++ks) {
      const double bf = Bs[warp * 8 + gq][ks * 4 + tq];
#pragma unroll
      for (int mi = 0; mi < 7; ++mi)
        if (mi < nmi) kb_dmma(p[mi][0], p[mi][1], __ldcg(&Xg[(size_t)(mi * 8 + gq) * ld + ks * 4 + tq]), bf);
    }
#pragma unroll
    for (int mi = 0; mi < 7; ++mi)
      if (mi < nmi) {
        double2 v = *reinterpret_cast<double2*>(&Cs[mi * 8 + gq][warp * 8 + 2 * tq]);
        v.x -= p[mi][0];
        v.y -= p[mi][1];
        *reinterpret_cast<double2*>(&Cs[mi * 8 + gq][warp * 8 + 2 * tq]) = v;
      }
  }
  __syncthreads();
  kb_load_tile_async(Bs, Wg, KB_TILE);
  kb_cp_async_commit();
  kb_cp_async_wait<0>();
  __syncthreads();
  double dl[7][2];
#pragma unroll
  for (int mi = 0; mi < 7; ++mi) dl[mi][0] = dl[mi][1] = 0.0;
  for (int ks = 0; ks < kmax; ++ks) {
    const double bf = Bs[warp * 8 + gq][ks * 4 + tq];
#pragma unroll
    for (int mi = 0; mi < 7; ++mi)
      if (mi < nmi) kb_dmma(dl[mi][0], dl[mi][1], Cs[mi * 8 + gq][ks * 4 + tq], bf);
  }
#pragma unroll
  for (int mi = 0; mi < 7; ++mi)
    if (mi < nmi)
      *reinterpret_cast<double2*>(&Xg[(size_t)(mi * 8 + gq) * ld + warp * 8 + 2 * tq]) =
          make_double2(o[mi][0] + dl[mi][0], o[mi][1] + dl[mi][1]);
}

// pivots of a diagonal tile spanning more than this ratio (as max L_ii / min L_ii) switch its column to
// kb_refined_solve.  tools/emulate_blocking.py sweep: any threshold from 10 to 1000 selects the same tiles.
#define KB_REFINE_RATIO 100.0
#ifdef KB_NO_REFINE            // A/B build aid: compile the corrected solves out
#define KB_REFINE_ON 0
#else
#define KB_REFINE_ON 1
#endif

struct KbCholArgs {
  KbAugGeom g;
  double* aug;
  double* dinv;
  int* prog;          // [B][NT]  number of final leading tile columns of each tile row
  int* pre;           // [B][nb]  1 once the PRE task of tile row r has written C1, C2
  int* pre_a;         // [B][nb]  1 once the early part (tile columns < r-2) of that PRE task is in
  int* ctl;           // [0] bulk queue head, [1] chain-SM tickets, [2] finished chains, [3] helper tickets,
                      // [8 + smid] CTA arrivals per SM, [8 + 512 + smid] chain-SM id + 1 or -1
  int* cstate;        // [B]  chain state of each candidate: 2 k = step k is next, 2 k + 1 = a CTA is running step k
  int helpers;        // bulk CTAs lend a hand with chain steps at both ends of the launch (kb_chain_helper)
  const int4* tasks;
  int ntasks;
  int B;
  int chain_sms;      // SMs reserved for chain workers
  int chain_wps;      // chain workers per reserved SM (1 or 2)
  int* info;
  int force_refine;   // debug aid KB_CHOL_REFINE: 1 = residual-corrected solves everywhere, -1 = nowhere (0: by pivot spread)
  unsigned long long* trace;
};

// ---------------------------------------------------------------------------
// Chain worker: critical path of its home candidates, round robin, never blocks on a step
// that is not ready.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void kb_chain_step(const KbCholArgs& P, KbCholSmem& sm, int b, int k,
                                              unsigned long long* stamps) {
  double* raw = &sm.ring[0];
  double (*C2)[KB_LDC] = reinterpret_cast<double (*)[KB_LDC]>(raw);
  double (*C1)[KB_LDC] = reinterpret_cast<double (*)[KB_LDC]>(raw + KB_TILE * KB_LDC);
  double (*Dk)[KB_LDC] = reinterpret_cast<double (*)[KB_LDC]>(raw + 2 * KB_TILE * KB_LDC);
  double* xch = raw + 3 * KB_TILE * KB_LDC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3, wr = warp >> 1, wc = warp & 1;
  const int ld = P.g.ld;
  double* A = P.aug + (size_t)b * P.g.stride;
  int* fl = P.prog + (size_t)b * P.g.NT;
  double* Tkk = A + (size_t)k * KB_TILE * ld + (size_t)k * KB_TILE;

  kb_load_tile_async(C2, Tkk, ld);
  if (k > 0) {
    kb_load_tile_async(C1, Tkk - KB_TILE, ld);
    kb_load_tile_async(Dk, P.dinv + ((size_t)b * P.g.nb + (k - 1)) * (KB_TILE * KB_TILE), KB_TILE);
  }
  kb_cp_async_commit();
  kb_cp_async_wait<0>();
  __syncthreads();
  if (stamps && tid == 0) stamps[1] = kb_globaltimer();
  if (k > 0) {
    // L[k,k-1] = C1 Dinv_{k-1}^T  -> shared (in place) and global, published at once
    double acc[4][4][2];
    double* Tsub = Tkk - KB_TILE;
    if (KB_REFINE_ON && Dk[0][KB_TILE - 1] != 0.0) {     // flag planted in the stored inverse by the previous step
      // ill-conditioned diagonal tile k-1: residual-corrected solve (writes shared C1 and global Tsub)
      kb_refined_solve<2>(C1, Dk, P.dinv + ((size_t)b * P.g.nb + (k - 1)) * (KB_TILE * KB_TILE),
                          Tkk - (size_t)KB_TILE * ld - KB_TILE, ld, Tsub, ld, C1);
    } else {
      kb_tri_multiply<2>(C1, Dk, acc);
      __syncthreads();
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int r = wr * 16 + mi * 8 + gq, c = (2 * ni + wc) * 8 + 2 * tq;
          const double2 v = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
          *reinterpret_cast<double2*>(&C1[r][c]) = v;
          *reinterpret_cast<double2*>(&Tsub[(size_t)r * ld + c]) = v;
        }
    }
    __syncthreads();
    if (tid == 0) {
      kb_flag_release(&fl[k], k);      // st.release.gpu: cumulative over the CTA barrier above
      if (stamps) stamps[2] = kb_globaltimer();
    }
    // C2 -= L L^T on the lower 8x8 blocks (each thread owns its accumulator positions)
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    unsigned need = 0;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
        if (wc * 4 + ni <= wr * 2 + mi) need |= 1u << (mi * 4 + ni);
    if (need) {
#pragma unroll 4
      for (int ks = 0; ks < KB_TILE / 4; ++ks) {
        double af[2], bf[4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) af[mi] = C1[wr * 16 + mi * 8 + gq][ks * 4 + tq];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) bf[ni] = C1[wc * 32 + ni * 8 + gq][ks * 4 + tq];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
            if (need & (1u << (mi * 4 + ni))) kb_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          if (need & (1u << (mi * 4 + ni))) {
            const int r = wr * 16 + mi * 8 + gq, c = wc * 32 + ni * 8 + 2 * tq;
            double2 v = *reinterpret_cast<double2*>(&C2[r][c]);
            v.x -= acc[mi][ni][0];
            v.y -= acc[mi][ni][1];
            *reinterpret_cast<double2*>(&C2[r][c]) = v;
          }
    }
    __syncthreads();
  }
  if (stamps && tid == 0) stamps[3] = kb_globaltimer();
  // Seated right-hand sides (KbAugGeom::seated): in the LAST tile rows nv .. nv+q-1 are [y F]^T, already
  // carried through every column (they are Z there) and updated by the product above.  Only the leading
  // nv x nv block is factored; the rows below get the triangular solve and the q x q corner becomes the
  // Schur block -Z Z^T the GLS kernel reads.  C1 (no longer needed) holds the saved rows and Z.
  const bool seated_last = P.g.seated && k == P.g.nb - 1;
  const int nv = P.g.n - k * KB_TILE;
  double (*Yb)[KB_LDC] = C1;
  if (seated_last) {
    for (int e = tid; e < 16 * KB_TILE; e += KB_THREADS) {
      const int j = e >> 6, c = e & 63, r = nv + j;
      double v = 0.0;
      if (r < KB_TILE) {
        v = C2[r][c];
        C2[r][c] = (r == c) ? 1.0 : 0.0;          // the factor routine sees identity padding, as always
      }
      Yb[j][c] = v;
    }
    for (int e = tid; e < nv * (KB_TILE - nv); e += KB_THREADS) C2[e / (KB_TILE - nv)][nv + e % (KB_TILE - nv)] = 0.0;
    __syncthreads();
  }
  kb_diag_factor_inverse(C2, Dk, xch, &P.info[b], k * KB_TILE);
  if (seated_last) {
    __syncthreads();
    // Z = Y Linv^T, one step of residual correction (as kb_refined_solve; always: 16 rows cost nothing):
    //   r = Y - Z L^T,  Z += r Linv^T
    // as three 16 x 64 x 64 triangular DMMA products -- warp w owns column block w and stops at the diagonal, as in the
    // thin tile tasks (scalar dot-product loops here used to make the last chain step 12 us longer than the others).
    // Yb rows 0..15 = Y, 16..31 = Z, 32..47 = r; columns >= nv of Z and r are kept at zero.
    const int ncb = (nv + 7) >> 3, kmaxw = 2 * (warp + 1), c0 = warp * 8 + 2 * tq;
    if (warp < ncb) {
      double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
      for (int ks = 0; ks < kmaxw; ++ks) {
        const double bf = Dk[warp * 8 + gq][ks * 4 + tq];
        kb_dmma(o[0][0], o[0][1], Yb[gq][ks * 4 + tq], bf);
        kb_dmma(o[1][0], o[1][1], Yb[8 + gq][ks * 4 + tq], bf);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
        *reinterpret_cast<double2*>(&Yb[16 + mi * 8 + gq][c0]) = make_double2(o[mi][0], o[mi][1]);
    }
    __syncthreads();
    if (warp < ncb) {
      double pr[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
      const int row = warp * 8 + gq;
      for (int ks = 0; ks < kmaxw; ++ks) {
        const int col = ks * 4 + tq;
        const double bf = (col <= row) ? C2[row][col] : 0.0;      // only the lower triangle of the factor is valid
        kb_dmma(pr[0][0], pr[0][1], Yb[16 + gq][col], bf);
        kb_dmma(pr[1][0], pr[1][1], Yb[24 + gq][col], bf);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const double2 yv = *reinterpret_cast<const double2*>(&Yb[mi * 8 + gq][c0]);
        *reinterpret_cast<double2*>(&Yb[32 + mi * 8 + gq][c0]) =
            make_double2(c0 < nv ? yv.x - pr[mi][0] : 0.0, c0 + 1 < nv ? yv.y - pr[mi][1] : 0.0);
      }
    }
    __syncthreads();
    if (warp < ncb) {
      double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
      for (int ks = 0; ks < kmaxw; ++ks) {
        const double bf = Dk[warp * 8 + gq][ks * 4 + tq];
        kb_dmma(o[0][0], o[0][1], Yb[32 + gq][ks * 4 + tq], bf);
        kb_dmma(o[1][0], o[1][1], Yb[40 + gq][ks * 4 + tq], bf);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const double2 zv = *reinterpret_cast<const double2*>(&Yb[16 + mi * 8 + gq][c0]);
        *reinterpret_cast<double2*>(&Yb[16 + mi * 8 + gq][c0]) =
            make_double2(c0 < nv ? zv.x + o[mi][0] : 0.0, c0 + 1 < nv ? zv.y + o[mi][1] : 0.0);
      }
    }
    __syncthreads();
    {
      const int a = tid >> 4, c = tid & 15;
      if (c <= a && nv + a < KB_TILE) {
        double s0 = 0.0, s1 = 0.0;
        int i = 0;
        for (; i + 1 < nv; i += 2) {
          s0 = fma(Yb[16 + a][i], Yb[16 + c][i], s0);
          s1 = fma(Yb[16 + a][i + 1], Yb[16 + c][i + 1], s1);
        }
        if (i < nv) s0 = fma(Yb[16 + a][i], Yb[16 + c][i], s0);
        C2[nv + a][nv + c] = Yb[a][nv + c] - (s0 + s1);
      }
    }
    for (int e = tid; e < 16 * nv; e += KB_THREADS) {
      const int j = e / nv, i = e - j * nv;
      if (nv + j < KB_TILE) C2[nv + j][i] = Yb[16 + j][i];
    }
    __syncthreads();
  }
  if (stamps && tid == 0) stamps[6] = kb_globaltimer();
  // pivot spread of this tile (its real rows only): beyond KB_REFINE_RATIO the panel solves of column k take the
  // residual-corrected path.  Every lane of warp 0 ends up with the flag; lane 31 plants it in the stored inverse.
  int rf_tile = 0;
  if (warp == 0) {
    const int nreal = (P.g.n - k * KB_TILE < KB_TILE) ? (P.g.n - k * KB_TILE) : KB_TILE;
    double lo = __longlong_as_double(0x7ff0000000000000LL), hi = 0.0;
    for (int i = lane; i < nreal; i += 32) {
      const double v = C2[i][i];
      lo = fmin(lo, v);
      hi = fmax(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    rf_tile = (P.force_refine > 0 || (P.force_refine == 0 && hi > KB_REFINE_RATIO * lo)) ? 1 : 0;
  }
  double* Dg = P.dinv + ((size_t)b * P.g.nb + k) * (KB_TILE * KB_TILE);
  for (int e = tid; e < KB_TILE * KB_TILE / 2; e += KB_THREADS) {
    const int r = e >> 5, c = (e & 31) * 2;
    double2 l = *reinterpret_cast<double2*>(&C2[r][c]);
    if (c > r) l.x = 0.0;
    if (c + 1 > r) l.y = 0.0;
    *reinterpret_cast<double2*>(&Tkk[(size_t)r * ld + c]) = l;
    double2 w = *reinterpret_cast<double2*>(&Dk[r][c]);
    // element [0][63] of the stored inverse (above the diagonal: never an operand) carries the refinement flag
    // of this tile to the bulk tasks, which stage the tile through shared memory anyway
    if (r == 0 && c == KB_TILE - 2) w.y = (double)rf_tile;      // (thread 31: warp 0)
    *reinterpret_cast<double2*>(&Dg[r * KB_TILE + c]) = w;
  }
  __syncthreads();
  if (tid == 0) kb_flag_release(&fl[k], k + 1);
}

// One chain step of candidate b if it is due and nobody else is on it.  With helpers on, the state of a candidate's
// chain lives in ONE global word, cstate[b] = 2 k (step k is next, nobody on it) or 2 k + 1 (a CTA is running step k),
// claimed with a compare-and-swap, so that any CTA may run a step: the dedicated chain workers walk their home
// candidates, and bulk CTAs with nothing else to do help at both ends of the launch (kb_chain_helper), where the
// chains, not the tile pool, are the critical path.  Without helpers (small populations: every candidate has a worker
// of its own and the launch is bound by the chain steps' latency) the owner keeps the step in shared memory (khint >= 0
// is then the step itself) and a step costs one flag round trip, as it always did.
// kmax: only steps below it (the start-up helpers stop after step 1).  Returns 1 if a step was run, 0 otherwise;
// *kdone = the candidate's next step afterwards, as far as this CTA knows.
__device__ __noinline__ int kb_chain_try(const KbCholArgs& P, KbCholSmem& sm, int b, int kmax, int khint, int* kdone) {
  const int tid = threadIdx.x, nb = P.g.nb;
  const bool locked = (P.helpers != 0);
  __syncthreads();
  if (tid == 0) {
    int k = khint, go = 0;
    if (locked) {
      // the state word and the PRE flag of the step we expect travel together (one round trip when the hint holds)
      const int hint = khint < nb ? khint : nb - 1;
      const int pre_hint = (hint <= 1) ? 1 : kb_flag_acquire(&P.pre[(size_t)b * nb + hint]);
      const int st = kb_flag_acquire(&P.cstate[b]);
      k = st >> 1;
      if (!(st & 1) && k < nb && k < kmax) {
        const int pre_ok = (k <= 1) ? 1 : (k == hint ? pre_hint : kb_flag_acquire(&P.pre[(size_t)b * nb + k]));
        if (pre_ok != 0 && atomicCAS(&P.cstate[b], 2 * k, 2 * k + 1) == 2 * k) {
          __threadfence();                                 // acquire side of the previous holder's release
          go = 1;
        }
      }
    } else if (k < nb && k < kmax) {
      go = (k <= 1) ? 1 : (kb_flag_acquire(&P.pre[(size_t)b * nb + k]) != 0);
    }
    sm.misc[2] = k;
    sm.ready = go;
  }
  __syncthreads();
  const int k = sm.misc[2];
  if (kdone) *kdone = k;
  if (sm.ready == 0) return 0;
  unsigned long long tr0 = 0;
  if (P.trace && tid == 0) tr0 = kb_globaltimer();
  unsigned long long* trec = P.trace ? P.trace + ((size_t)P.ntasks + (size_t)b * nb + k) * 8 : nullptr;
  kb_chain_step(P, sm, b, k, trec);
  if (P.trace && tid == 0) {
    // chain record: start | tiles loaded | sub-diagonal tile published | factor start | end | meta | factor end
    trec[0] = tr0; trec[4] = kb_globaltimer();
    trec[5] = ((unsigned long long)b << 40) | ((unsigned long long)k << 24) | ((unsigned long long)k << 8) | 0;
    trec[7] = blockIdx.x;
  }
  __syncthreads();
  if (tid == 0) {
    if (locked) kb_flag_release(&P.cstate[b], 2 * (k + 1));        // cumulative over the CTA barrier above
    if (k + 1 == nb) atomicAdd(&P.ctl[2], 1);
  }
  if (kdone) *kdone = k + 1;
  return 1;
}

// ---------------------------------------------------------------------------
// Chain worker: critical path of its home candidates, round robin, never blocks on a step
// that is not ready.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void kb_chain_worker(const KbCholArgs& P, KbCholSmem& sm, int wid, int nworkers) {
  // home candidates b = wid + j * nworkers (at most 8: the host sizes chain_sms for it); the next step of each, as
  // far as this worker knows, lives in shared memory and the loop stays ROLLED: one copy of the (large) step in the
  // instruction stream
  const int tid = threadIdx.x, nb = P.g.nb;
  if (tid < 8) sm.knext[tid] = 0;
  __syncthreads();
  while (true) {
    bool progressed = false, all_done = true;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      const int b = wid + j * nworkers;
      if (b >= P.B) break;
      const int kh = sm.knext[j];
      if (kh >= nb) continue;
      int kd = kh;
      if (kb_chain_try(P, sm, b, 0x7fffffff, kh, &kd)) progressed = true;
      if (tid == 0) sm.knext[j] = kd;
      if (kd < nb) all_done = false;
    }
    __syncthreads();
    if (all_done) break;
    if (!progressed) __nanosleep(200);
  }
}

// Bulk CTA lending a hand.  phase 0 (start of the launch, before its first tile task): steps 0 and 1 of ONE candidate --
// they depend on nothing, and 2 B of them on the few chain workers used to keep the whole tile pool waiting for its
// first inverted diagonal blocks.  phase 1 (tile queue exhausted): any step of any candidate until every chain is done;
// by then the chains are all that is left and the dedicated workers (two per SM, sharing one FP64 pipe) are the slow
// way to run them.
__device__ __forceinline__ void kb_chain_helper(const KbCholArgs& P, KbCholSmem& sm, int phase) {
  const int tid = threadIdx.x;
  if (phase == 0) {
    __syncthreads();
    if (tid == 0) sm.misc[3] = atomicAdd(&P.ctl[3], 1);
    __syncthreads();
    const int t = sm.misc[3];
    if (t >= P.B) return;
    // candidates from the far end first: the dedicated workers start with b = worker id
    const int b = P.B - 1 - t;
    if (kb_chain_try(P, sm, b, 2, 0, nullptr)) kb_chain_try(P, sm, b, 2, 1, nullptr);
    return;
  }
  int start = blockIdx.x % P.B;
  while (true) {
    __syncthreads();
    if (tid == 0) sm.misc[3] = kb_flag_acquire(&P.ctl[2]);
    __syncthreads();
    if (sm.misc[3] >= P.B) break;
    bool progressed = false;
#pragma unroll 1
    for (int i = 0; i < P.B; ++i) {
      int b = start + i;
      if (b >= P.B) b -= P.B;
      if (kb_chain_try(P, sm, b, 0x7fffffff, P.g.nb - 1, nullptr)) { progressed = true; start = b; break; }
    }
    if (!progressed) __nanosleep(300);
  }
}

// ---------------------------------------------------------------------------
// Bulk K loop.  MODE 0: 64-row tile task (A 64 rows, B 64 rows, 16 x 32 per warp)
//               MODE 1: PRE task: A = tile row r, B = tile row r-1 (acc[0..1]); second
//                       accumulator acc[2..3] with B = A (lower 8x8 blocks only)
//               MODE 2: thin task: A = 16 rows, B 64 rows, warp w owns column block w
//               MODE 4: 128-row task (two tile rows i, i+1 against one B row): 32 x 32 per warp,
//                       16 DMMA per 8 fragment loads, two pipeline stages
//               MODE 5: short task: the LAST tile row of R when it holds at most 56 live rows (n mod 64 sites plus
//                       the seated right-hand sides): A = the first 8 * need2 rows only, warp w owns column block w
//                       (need2 DMMA per need2 + 1 fragment loads).  n = 2000: 24 of 64 rows, n = 1000: 48 -- the
//                       padding rows of that tile row were 5 % / 3 % of the whole factorisation.
// ---------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void kb_kloop(KbCholSmem& sm, const double* __restrict__ Arow,
                                         const double* __restrict__ Brow, bool same, int ld, int k0, int k1,
                                         const int* fa, const int* fa2, const int* fb, int ready_cols,
                                         unsigned need2, double (&acc)[4][4][2]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3, wr = warp >> 1, wc = warp & 1;
  constexpr int KC = (MODE == 4) ? KB_KC128 : KB_KC64;     // chunk width
  constexpr int LDK = KC + 4;                              // padded row stride of a chunk row
  constexpr int CPT = KB_TILE / KC;                        // chunks per tile column
  constexpr int NS = (MODE == 4) ? KB_NS128 : KB_NS64;     // pipeline stages
  constexpr int AROWS = (MODE == 4) ? 128 : 64;
  constexpr int SROWS = AROWS + 64;                        // ring rows per stage
  constexpr int UPR = KC / 2;                              // 16-byte units per chunk row
  const int c_begin = k0 * CPT, c_end = k1 * CPT;

  auto load_chunk = [&](int c, int stage) {
    const int col0 = c * KC;
    double* as = &sm.ring[stage * SROWS * LDK];
    double* bs = as + AROWS * LDK;
    if (MODE == 2) {
      if (tid < 16 * UPR) {
        const int row = tid / UPR, seg = (tid % UPR) * 2;
        kb_cp_async16(&as[row * LDK + seg], Arow + (size_t)row * ld + col0 + seg);
      }
    } else if (MODE == 5) {
      for (int idx = tid; idx < (int)need2 * 8 * UPR; idx += KB_THREADS) {
        const int row = idx / UPR, seg = (idx % UPR) * 2;
        kb_cp_async16(&as[row * LDK + seg], Arow + (size_t)row * ld + col0 + seg);
      }
    } else {
#pragma unroll
      for (int j = 0; j < (AROWS * UPR) / KB_THREADS; ++j) {
        const int idx = tid + j * KB_THREADS;
        const int row = idx / UPR, seg = (idx % UPR) * 2;
        kb_cp_async16(&as[row * LDK + seg], Arow + (size_t)row * ld + col0 + seg);
      }
    }
    if (!same) {
#pragma unroll
      for (int j = 0; j < (KB_TILE * UPR) / KB_THREADS; ++j) {
        const int idx = tid + j * KB_THREADS;
        const int row = idx / UPR, seg = (idx % UPR) * 2;
        kb_cp_async16(&bs[row * LDK + seg], Brow + (size_t)row * ld + col0 + seg);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < NS - 1; ++s) {
    if (c_begin + s < c_end) load_chunk(c_begin + s, s);
    kb_cp_async_commit();
  }
  for (int c = c_begin; c < c_end; ++c) {
    const int cn = c + NS - 1;
    const int pn = cn / CPT;                  // tile column the next prefetch reads
    const bool poll = (cn < c_end) && (pn >= ready_cols);
    if (poll && tid == 0) {
      int ra, rb, rc;
      const unsigned long long tp0 = kb_globaltimer();
      while (true) {
        ra = kb_flag_acquire(fa);
        rb = kb_flag_acquire(fb);
        rc = kb_flag_acquire(fa2);
        if (ra > pn && rb > pn && rc > pn) break;
        __nanosleep(64);
      }
      sm.poll_ns += kb_globaltimer() - tp0;
      ra = ra < rb ? ra : rb;
      sm.ready = ra < rc ? ra : rc;
    }
    kb_cp_async_wait<NS - 2>();
    __syncthreads();
    if (poll) ready_cols = sm.ready;
    // the stage freed by the barrier is refilled after the first group of products is queued,
    // so the copy instructions issue in the shadow of the DMMAs instead of in front of them
    auto refill = [&]() {
      if (cn < c_end) load_chunk(cn, (cn - c_begin) % NS);
      kb_cp_async_commit();
    };
    const int stg = (c - c_begin) % NS;
    const double (*as)[LDK] = reinterpret_cast<const double (*)[LDK]>(&sm.ring[stg * SROWS * LDK]);
    const double (*bs)[LDK] = same ? as : as + AROWS;
    if (MODE == 2) {
#pragma unroll
      for (int ks = 0; ks < KC / 4; ++ks) {
        const double a0 = as[gq][ks * 4 + tq], a1 = as[8 + gq][ks * 4 + tq];
        const double bf = bs[warp * 8 + gq][ks * 4 + tq];
        kb_dmma(acc[0][0][0], acc[0][0][1], a0, bf);
        kb_dmma(acc[1][0][0], acc[1][0][1], a1, bf);
        if (ks == 0) refill();
      }
    } else if (MODE == 5) {
#pragma unroll
      for (int ks = 0; ks < KC / 4; ++ks) {
        const double bf = bs[warp * 8 + gq][ks * 4 + tq];
        double af[7];
#pragma unroll
        for (int mi = 0; mi < 7; ++mi) af[mi] = as[mi * 8 + gq][ks * 4 + tq];      // (rows past the live ones: stale, unused)
#pragma unroll
        for (int mi = 0; mi < 7; ++mi)
          if (mi < (int)need2) kb_dmma(acc[mi & 3][mi >> 2][0], acc[mi & 3][mi >> 2][1], af[mi], bf);
        if (ks == 0) refill();
      }
    } else if (MODE == 4) {
#pragma unroll
      for (int ks = 0; ks < KC / 4; ++ks) {
        double af[4], bf[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) af[mi] = as[wr * 32 + mi * 8 + gq][ks * 4 + tq];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) bf[ni] = bs[wc * 32 + ni * 8 + gq][ks * 4 + tq];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) kb_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        if (ks == 0) refill();
      }
    } else {
#pragma unroll
      for (int ks = 0; ks < KC / 4; ++ks) {
        double af[2], bf[4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) af[mi] = as[wr * 16 + mi * 8 + gq][ks * 4 + tq];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) bf[ni] = bs[wc * 32 + ni * 8 + gq][ks * 4 + tq];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) kb_dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        if (MODE == 1 && need2) {
          double b2[4];
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) b2[ni] = as[wc * 32 + ni * 8 + gq][ks * 4 + tq];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
              if (need2 & (1u << (mi * 4 + ni))) kb_dmma(acc[2 + mi][ni][0], acc[2 + mi][ni][1], af[mi], b2[ni]);
        }
        if (ks == 0) refill();
      }
    }
  }
  kb_cp_async_wait<0>();
  __syncthreads();
}

__device__ __forceinline__ int kb_wait_rows(KbCholSmem& sm, const int* fa, const int* fa2, const int* fb, int need) {
  if (threadIdx.x == 0) {
    int ra, rb, rc;
    while (true) {
      ra = kb_flag_acquire(fa);
      rb = kb_flag_acquire(fb);
      rc = kb_flag_acquire(fa2);
      if (ra >= need && rb >= need && rc >= need) break;
      __nanosleep(64);
    }
    ra = ra < rb ? ra : rb;
    sm.ready = ra < rc ? ra : rc;
  }
  __syncthreads();
  return sm.ready;
}

// wait for the chain's Dinv_k and start staging it behind the C tile(s); the copy overlaps the
// read of the tile being updated.  kb_dinv_land() completes it.
__device__ __forceinline__ void kb_dinv_issue(const KbCholArgs& P, KbCholSmem& sm, double (*Ds)[KB_LDC], const int* flk,
                                              int b, int tkc) {
  if (threadIdx.x == 0)
    while (kb_flag_acquire(flk) < tkc + 1) __nanosleep(64);
  __syncthreads();
  kb_load_tile_async(Ds, P.dinv + ((size_t)b * P.g.nb + tkc) * (KB_TILE * KB_TILE), KB_TILE);
  kb_cp_async_commit();
}
__device__ __forceinline__ void kb_dinv_land() {
  kb_cp_async_wait<0>();
  __syncthreads();
}

__device__ __forceinline__ void kb_bulk_worker(const KbCholArgs& P, KbCholSmem& sm) {
  double (*Cs)[KB_LDC] = reinterpret_cast<double (*)[KB_LDC]>(&sm.ring[0]);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3, wr = warp >> 1, wc = warp & 1;
  const int NT = P.g.NT, ld = P.g.ld, nb = P.g.nb;

  // (Claiming the next ticket early, during the current K loop, was measured 10 % SLOWER: a claimed
  // but unstarted task delays the urgent head-of-column tasks behind a long K loop.)
  if (tid == 32) {
    const int nx = atomicAdd(&P.ctl[0], 1);
    sm.task = nx;
    if (nx < P.ntasks) sm.tdesc = P.tasks[nx];
  }
  while (true) {
    __syncthreads();
    const int tk = sm.task;
    if (tk >= P.ntasks) break;
    const int4 T = sm.tdesc;
    const int b = T.x, ti = T.y, tkc = T.z, type = T.w & 0xff, kstart = T.w >> 8;
    unsigned long long tr0 = 0, tr1 = 0, tr2 = 0, tr3 = 0;
    if (P.trace && tid == 0) tr0 = kb_globaltimer();
    if (tid == 0) sm.poll_ns = 0;
    double* A = P.aug + (size_t)b * P.g.stride;
    int* fl = P.prog + (size_t)b * NT;
    const bool is_pre = (type == KB_TASK_PRE || type == KB_TASK_PRE_A);
    const bool is_schur = (type == KB_TASK_SCHUR || type == KB_TASK_SCHUR_THIN || type == KB_TASK_SCHUR_THIN_A);
    const bool thin = (type == KB_TASK_THIN || type == KB_TASK_SCHUR_THIN || type == KB_TASK_SCHUR_THIN_A);
    const bool two = (type == KB_TASK_OFF2);
    const bool shortrow = (type == KB_TASK_OFFS);
    const int nmi = (P.g.n - (nb - 1) * KB_TILE + (P.g.seated ? P.g.q : 0) + 7) >> 3;    // live 8-row blocks of the last tile row
    const int brow = is_pre ? ti - 1 : tkc;
    const int kend = (type == KB_TASK_PRE_A) ? ti - 2
                     : (is_pre ? ti - 1 : (type == KB_TASK_SCHUR_THIN_A ? nb - 1 : (tkc < nb ? tkc : nb)));
    const double* Arow = A + (size_t)ti * KB_TILE * ld;
    const double* Brow = A + (size_t)brow * KB_TILE * ld;
    double* Atile = A + (size_t)ti * KB_TILE * ld + (size_t)(is_pre ? ti - 1 : tkc) * KB_TILE;
    const int* fa = &fl[ti];
    const int* fa2 = two ? &fl[ti + 1] : fa;
    const int* fb = &fl[brow];
    // the tile(s) this task updates are read once, after the K loop: pull them towards L2 now
    {
      const int nrows = thin ? 16 : (two ? 128 : (shortrow ? nmi * 8 : 64));
      const int lines = (is_pre ? 8 : 4);          // 128-byte lines per row (PRE touches two tiles)
      for (int e = tid; e < nrows * lines; e += KB_THREADS) {
        const double* p = Atile + (size_t)(e / lines) * ld + (e % lines) * 16;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      }
    }
    int ready_cols = 0;
    if (type == KB_TASK_PRE && kstart > 0) {       // the early part of this PRE task must be in
      if (tid == 0)
        while (kb_flag_acquire(&P.pre_a[(size_t)b * nb + ti]) == 0) __nanosleep(64);
      __syncthreads();
    }
    if (type == KB_TASK_SCHUR_THIN && kstart > 0) {   // ... and so must the early part of the Schur sum
      if (tid == 0)
        while (kb_flag_acquire(&P.pre_a[(size_t)b * nb]) == 0) __nanosleep(64);
      __syncthreads();
    }
    if (kend > kstart) ready_cols = kb_wait_rows(sm, fa, fa2, fb, kstart + 1);
    if (P.trace && tid == 0) tr1 = kb_globaltimer();

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    unsigned need2 = 0;
    if (is_pre) {
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          if (wc * 4 + ni <= wr * 2 + mi) need2 |= 1u << (mi * 4 + ni);
      kb_kloop<1>(sm, Arow, Brow, false, ld, kstart, kend, fa, fa2, fb, ready_cols, need2, acc);
    } else if (thin) {
      kb_kloop<2>(sm, Arow, Brow, false, ld, kstart, kend, fa, fa2, fb, ready_cols, 0u, acc);
    } else if (two) {
      kb_kloop<4>(sm, Arow, Brow, false, ld, kstart, kend, fa, fa2, fb, ready_cols, 0u, acc);
    } else if (shortrow) {
      kb_kloop<5>(sm, Arow, Brow, false, ld, kstart, kend, fa, fa2, fb, ready_cols, (unsigned)nmi, acc);
    } else {
      kb_kloop<0>(sm, Arow, Brow, ti == tkc, ld, kstart, kend, fa, fa2, fb, ready_cols, 0u, acc);
    }
    if (P.trace && tid == 0) tr2 = kb_globaltimer();
#if KB_EARLY_CLAIM
    // the next ticket is claimed as the K loop ends (every thread copied this task's descriptor long ago), so the
    // atomic's and the descriptor's round trips run under the epilogue
    if (tid == 32) {
      const int nx = atomicAdd(&P.ctl[0], 1);
      sm.task = nx;
      if (nx < P.ntasks) sm.tdesc = P.tasks[nx];
    }
#endif

    if (is_pre) {
      // C1 = A[r,r-1] - acc (full tile), C2 = A[r,r] - acc2 (lower 8x8 blocks), both in place
      double* T1 = Atile;
      double* T2 = T1 + KB_TILE;
      // all loads first (the compiler cannot hoist them over possibly aliasing stores), then all stores
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int r = wr * 16 + mi * 8 + gq, c = wc * 32 + ni * 8 + 2 * tq;
          const double2 v = __ldcg(reinterpret_cast<const double2*>(&T1[(size_t)r * ld + c]));
          acc[mi][ni][0] = v.x - acc[mi][ni][0];
          acc[mi][ni][1] = v.y - acc[mi][ni][1];
          if (need2 & (1u << (mi * 4 + ni))) {
            const double2 u = __ldcg(reinterpret_cast<const double2*>(&T2[(size_t)r * ld + c]));
            acc[2 + mi][ni][0] = u.x - acc[2 + mi][ni][0];
            acc[2 + mi][ni][1] = u.y - acc[2 + mi][ni][1];
          }
        }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int r = wr * 16 + mi * 8 + gq, c = wc * 32 + ni * 8 + 2 * tq;
          *reinterpret_cast<double2*>(&T1[(size_t)r * ld + c]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
          if (need2 & (1u << (mi * 4 + ni)))
            *reinterpret_cast<double2*>(&T2[(size_t)r * ld + c]) = make_double2(acc[2 + mi][ni][0], acc[2 + mi][ni][1]);
        }
    } else if (thin) {
      // 16 x 64 strip; warp w owns column block w
      double c0v[2][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const double2 v = __ldcg(reinterpret_cast<const double2*>(&Atile[(size_t)(mi * 8 + gq) * ld + warp * 8 + 2 * tq]));
        c0v[mi][0] = v.x - acc[mi][0][0];
        c0v[mi][1] = v.y - acc[mi][0][1];
      }
      if (is_schur) {
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
          *reinterpret_cast<double2*>(&Atile[(size_t)(mi * 8 + gq) * ld + warp * 8 + 2 * tq]) =
              make_double2(c0v[mi][0], c0v[mi][1]);
      } else {
        double (*Ds)[KB_LDC] = Cs + KB_TILE;
        kb_dinv_issue(P, sm, Ds, &fl[tkc], b, tkc);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
          *reinterpret_cast<double2*>(&Cs[mi * 8 + gq][warp * 8 + 2 * tq]) = make_double2(c0v[mi][0], c0v[mi][1]);
        kb_dinv_land();
        if (P.trace && tid == 0) tr3 = kb_globaltimer();
        if (KB_REFINE_ON && Ds[0][KB_TILE - 1] != 0.0) {
          kb_refined_solve_thin(Cs, Ds, P.dinv + ((size_t)b * nb + tkc) * (KB_TILE * KB_TILE),
                                A + (size_t)tkc * KB_TILE * ld + (size_t)tkc * KB_TILE, ld, Atile, ld);
        } else {
          double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
          const int kmax = 2 * (warp + 1);
          for (int ks = 0; ks < kmax; ++ks) {
            const double bf = Ds[warp * 8 + gq][ks * 4 + tq];
            kb_dmma(o[0][0], o[0][1], Cs[gq][ks * 4 + tq], bf);
            kb_dmma(o[1][0], o[1][1], Cs[8 + gq][ks * 4 + tq], bf);
          }
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
            *reinterpret_cast<double2*>(&Atile[(size_t)(mi * 8 + gq) * ld + warp * 8 + 2 * tq]) =
                make_double2(o[mi][0], o[mi][1]);
        }
      }
    } else if (shortrow) {
      // 8 nmi x 64 strip of the last tile row; warp w owns column block w (the rows below stay the zeros the
      // correlation build wrote: identity padding)
      double (*Ds)[KB_LDC] = Cs + KB_TILE;
      kb_dinv_issue(P, sm, Ds, &fl[tkc], b, tkc);
#pragma unroll
      for (int mi = 0; mi < 7; ++mi)
        if (mi < nmi) {
          const double2 v = __ldcg(reinterpret_cast<const double2*>(&Atile[(size_t)(mi * 8 + gq) * ld + warp * 8 + 2 * tq]));
          *reinterpret_cast<double2*>(&Cs[mi * 8 + gq][warp * 8 + 2 * tq]) =
              make_double2(v.x - acc[mi & 3][mi >> 2][0], v.y - acc[mi & 3][mi >> 2][1]);
        }
      kb_dinv_land();
      if (P.trace && tid == 0) tr3 = kb_globaltimer();
      if (KB_REFINE_ON && Ds[0][KB_TILE - 1] != 0.0) {
        kb_refined_solve_short(Cs, Ds, P.dinv + ((size_t)b * nb + tkc) * (KB_TILE * KB_TILE),
                               A + (size_t)tkc * KB_TILE * ld + (size_t)tkc * KB_TILE, ld, Atile, ld, nmi);
      } else {
        double o[7][2];
#pragma unroll
        for (int mi = 0; mi < 7; ++mi) o[mi][0] = o[mi][1] = 0.0;
        const int kmax = 2 * (warp + 1);
        for (int ks = 0; ks < kmax; ++ks) {
          const double bf = Ds[warp * 8 + gq][ks * 4 + tq];
#pragma unroll
          for (int mi = 0; mi < 7; ++mi)
            if (mi < nmi) kb_dmma(o[mi][0], o[mi][1], Cs[mi * 8 + gq][ks * 4 + tq], bf);
        }
#pragma unroll
        for (int mi = 0; mi < 7; ++mi)
          if (mi < nmi)
            *reinterpret_cast<double2*>(&Atile[(size_t)(mi * 8 + gq) * ld + warp * 8 + 2 * tq]) = make_double2(o[mi][0], o[mi][1]);
      }
    } else if (two) {
      // 128 x 64: C = A[i..i+1, k] - acc -> shared, one triangular multiply for both tile rows
      double (*Ds)[KB_LDC] = Cs + 2 * KB_TILE;
      kb_dinv_issue(P, sm, Ds, &fl[tkc], b, tkc);
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int r = wr * 32 + mi * 8 + gq, c = wc * 32 + ni * 8 + 2 * tq;
          const double2 v = __ldcg(reinterpret_cast<const double2*>(&Atile[(size_t)r * ld + c]));
          *reinterpret_cast<double2*>(&Cs[r][c]) = make_double2(v.x - acc[mi][ni][0], v.y - acc[mi][ni][1]);
        }
      kb_dinv_land();
      if (P.trace && tid == 0) tr3 = kb_globaltimer();
      if (KB_REFINE_ON && Ds[0][KB_TILE - 1] != 0.0) {
        kb_refined_solve<4>(Cs, Ds, P.dinv + ((size_t)b * nb + tkc) * (KB_TILE * KB_TILE),
                            A + (size_t)tkc * KB_TILE * ld + (size_t)tkc * KB_TILE, ld, Atile, ld, nullptr);
      } else {
        kb_tri_multiply<4>(Cs, Ds, acc);
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const int r = wr * 32 + mi * 8 + gq, c = (2 * ni + wc) * 8 + 2 * tq;
            *reinterpret_cast<double2*>(&Atile[(size_t)r * ld + c]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
          }
      }
    } else {
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int r = wr * 16 + mi * 8 + gq, c = wc * 32 + ni * 8 + 2 * tq;
          const double2 v = __ldcg(reinterpret_cast<const double2*>(&Atile[(size_t)r * ld + c]));
          acc[mi][ni][0] = v.x - acc[mi][ni][0];
          acc[mi][ni][1] = v.y - acc[mi][ni][1];
        }
      if (is_schur) {
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const int r = wr * 16 + mi * 8 + gq, c = wc * 32 + ni * 8 + 2 * tq;
            *reinterpret_cast<double2*>(&Atile[(size_t)r * ld + c]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
          }
      } else {
        // L[i,k] = C * inv(L[k,k])^T : wait for the chain, then one triangular 64^3 DMMA multiply
        double (*Ds)[KB_LDC] = Cs + KB_TILE;
        kb_dinv_issue(P, sm, Ds, &fl[tkc], b, tkc);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const int r = wr * 16 + mi * 8 + gq, c = wc * 32 + ni * 8 + 2 * tq;
            *reinterpret_cast<double2*>(&Cs[r][c]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
          }
        kb_dinv_land();
        if (P.trace && tid == 0) tr3 = kb_globaltimer();
        if (KB_REFINE_ON && Ds[0][KB_TILE - 1] != 0.0) {
          kb_refined_solve<2>(Cs, Ds, P.dinv + ((size_t)b * nb + tkc) * (KB_TILE * KB_TILE),
                              A + (size_t)tkc * KB_TILE * ld + (size_t)tkc * KB_TILE, ld, Atile, ld, nullptr);
        } else {
          kb_tri_multiply<2>(Cs, Ds, acc);
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
              const int r = wr * 16 + mi * 8 + gq, c = (2 * ni + wc) * 8 + 2 * tq;
              *reinterpret_cast<double2*>(&Atile[(size_t)r * ld + c]) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
        }
      }
    }
    // ---- publish (st.release.gpu is cumulative over the CTA barrier) and, in parallel on another
    //      warp, claim the next ticket --------------------------------------------------------
    __syncthreads();
    if (tid == 0) {
      if (type == KB_TASK_PRE) kb_flag_release(&P.pre[(size_t)b * nb + ti], 1);
      else if (type == KB_TASK_PRE_A) kb_flag_release(&P.pre_a[(size_t)b * nb + ti], 1);
      else if (type == KB_TASK_SCHUR_THIN_A) kb_flag_release(&P.pre_a[(size_t)b * nb], 1);   // slot 0 is free
      else if (!is_schur) {
        kb_flag_release(&fl[ti], tkc + 1);
        if (two) kb_flag_release(&fl[ti + 1], tkc + 1);
      }
    }
#if !KB_EARLY_CLAIM
    if (tid == 32) {
      const int nx = atomicAdd(&P.ctl[0], 1);
      sm.task = nx;
      if (nx < P.ntasks) sm.tdesc = P.tasks[nx];
    }
#endif
    if (P.trace && tid == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      unsigned long long* t = P.trace + (size_t)tk * 8;
      t[0] = tr0; t[1] = tr1; t[2] = tr2; t[3] = tr3; t[4] = kb_globaltimer();
      t[5] = ((unsigned long long)b << 40) | ((unsigned long long)ti << 24) | ((unsigned long long)tkc << 8) | type;
      t[6] = smid | (sm.poll_ns << 16); t[7] = blockIdx.x;
    }
  }
}

__global__ void __launch_bounds__(KB_THREADS, 2)
kb_chol_kernel(KbCholArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  KbCholSmem& sm = *reinterpret_cast<KbCholSmem*>(smem_raw);
  const int tid = threadIdx.x;
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&P);
    if (tid < (int)((sizeof(KbCholArgs) + 7) / 8)) sm.argbuf[tid] = src[tid];
  }
  // ---- role: the first chain_sms SMs to start a CTA host the chain workers ----
  if (tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    int* arrivals = P.ctl + 8;
    int* smrole = P.ctl + 8 + 512;
    const int slot = atomicAdd(&arrivals[smid], 1);
    int role;
    if (slot == 0) {
      const int t = atomicAdd(&P.ctl[1], 1);
      role = (t < P.chain_sms) ? t + 1 : -1;
      kb_flag_release(&smrole[smid], role);
    } else {
      while ((role = kb_flag_acquire(&smrole[smid])) == 0) __nanosleep(32);
    }
    sm.misc[0] = role;
    sm.misc[1] = slot;
  }
  __syncthreads();
  const int role = sm.misc[0], slot = sm.misc[1];
  static_assert(sizeof(KbCholArgs) <= sizeof(sm.argbuf), "argument copy does not fit");
  KbCholArgs& Ps = *reinterpret_cast<KbCholArgs*>(sm.argbuf);
  if (role > 0) {
    if (slot < P.chain_wps) {
      kb_chain_worker(Ps, sm, (role - 1) * P.chain_wps + slot, P.chain_sms * P.chain_wps);
    } else {
      // spare CTA on a chain SM: stay off the FP64 pipe until every chain has finished
      if (tid == 0)
        while (kb_flag_acquire(&P.ctl[2]) < P.B) __nanosleep(500);
      __syncthreads();
    }
    kb_bulk_worker(P, sm);
    return;
  }
  // bulk SM: its first CTA helps the chains before its first and after its last tile task (one helper per SM: a
  // second one would share the FP64 pipe)
  const bool helper = P.helpers && slot == 0;
  if (helper && (P.helpers & 1)) kb_chain_helper(Ps, sm, 0);
  kb_bulk_worker(P, sm);
  if (helper && (P.helpers & 2)) kb_chain_helper(Ps, sm, 1);
}

cudaError_t kb_launch_chol(cudaStream_t st, const KbAugGeom& g, int B, double* aug, double* dinv,
                           int* sync_words, const int4* tasks, int ntasks, int* info, int num_sms) {
  const size_t smem = sizeof(KbCholSmem);
  cudaError_t e = cudaFuncSetAttribute(kb_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const size_t nwords = kb_chol_sync_words(g, B);
  e = cudaMemsetAsync(sync_words, 0, sizeof(int) * nwords, st);
  if (e != cudaSuccess) return e;
  KbCholArgs P{};
  P.g = g; P.aug = aug; P.dinv = dinv;
  P.ctl = sync_words;
  P.prog = sync_words + KB_CHOL_CTL_WORDS;
  P.pre = P.prog + (size_t)B * g.NT;
  P.pre_a = P.pre + (size_t)B * g.nb;
  P.cstate = P.pre_a + (size_t)B * g.nb;
  static const char* refine_env = getenv("KB_CHOL_REFINE");       // debug aid: 1 everywhere, 0 nowhere
  P.force_refine = refine_env ? (atoi(refine_env) > 0 ? 1 : -1) : 0;
  P.tasks = tasks; P.ntasks = ntasks; P.B = B; P.info = info;
  // chain SMs: up to 16; one worker per SM while that covers the population, else two; a
  // worker serves at most 8 candidates
  // (measured, B = 64: 16 SMs are best at 16 tile columns, 8 at 32 -- the bulk work per chain step
  // grows with the column count, so fewer chain SMs keep up)
  const int csm_cap = g.nb >= 24 ? 8 : (g.nb > 16 ? 12 : 16);
  int csm = B < csm_cap ? B : csm_cap;
  int wps = (B > csm) ? 2 : 1;
  while (csm * wps * 8 < B && csm < num_sms / 2) ++csm;
  static const char* csm_env = getenv("KB_CHOL_CHAIN_SMS");     // debug aids
  static const char* wps_env = getenv("KB_CHOL_CHAIN_WPS");
  if (csm_env && atoi(csm_env) > 0) csm = atoi(csm_env);
  if (wps_env && atoi(wps_env) > 0) wps = atoi(wps_env);
  if (csm * wps > B) csm = (B + wps - 1) / wps;
  while (csm * wps * 8 < B) ++csm;
  P.chain_sms = csm;
  P.chain_wps = wps;
  // helpers only where a chain worker serves several candidates: with a worker per candidate the launch is bound by
  // the latency of the chain steps themselves and the claim protocol would only add to it
  static const char* help_env = getenv("KB_CHOL_HELPERS");      // debug aid: bit 0 start-up helpers, bit 1 drain helpers
  P.helpers = help_env ? atoi(help_env) : ((B > csm * wps) ? KB_CHOL_HELPERS_DEFAULT : 0);
  int grid = num_sms * 2;
  static const char* grid_env = getenv("KB_CHOL_CTAS_PER_SM");   // debug aid: 1 or 2
  if (grid_env && atoi(grid_env) == 1) { grid = num_sms; P.chain_wps = 1; }
  // Debug aid: KB_CHOL_TRACE=<file> dumps per-task timestamps (ns) of this launch.
  static const char* trace_path = getenv("KB_CHOL_TRACE");
  const size_t nrec = (size_t)ntasks + (size_t)B * g.nb;
  P.trace = nullptr;
  if (trace_path) {
    e = cudaMalloc(&P.trace, sizeof(unsigned long long) * 8 * nrec);
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(P.trace, 0, sizeof(unsigned long long) * 8 * nrec, st);
  }
  // Cooperative launch: every CTA is resident, so the chain workers exist while bulk CTAs spin.
  void* params[] = {&P};
  e = cudaLaunchCooperativeKernel((const void*)kb_chol_kernel, dim3(grid), dim3(KB_THREADS), params, smem, st);
  if (e != cudaSuccess) return e;
  if (trace_path) {
    std::vector<unsigned long long> h(nrec * 8);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), P.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    FILE* f = fopen(trace_path, "wb");
    if (f) { fwrite(h.data(), sizeof(unsigned long long), h.size(), f); fclose(f); }
    cudaFree(P.trace);
  }
  return cudaSuccess;
}

size_t kb_chol_sync_words(const KbAugGeom& g, int B) {
  return (size_t)KB_CHOL_CTL_WORDS + (size_t)B * g.NT + 2 * (size_t)B * g.nb + (size_t)B;
}

// Bulk task queue: every dependency precedes its user.  Column k holds the tile tasks (i,k),
// i >= k+2 (the sub-diagonal tile belongs to the chain), with the tile the chain needs next
// (k+2,k) and the PRE task of row k+2 first.
// Candidate groups and their stagger.  Every candidate walks the same column sequence; with all of them in
// lock-step the first and the last columns of the factorisation (little bulk work per chain step) leave the bulk
// CTAs waiting for the chain workers.  The population is therefore dealt into `groups` groups and group j enters the
// queue `stagger` tile columns after group j-1: one group's thin columns overlap another group's wide ones.  Any
// interleaving keeps the queue's invariant (every dependency of a task is earlier in the queue), because tasks of
// different candidates are independent and each candidate's own order is unchanged.
static void kb_chol_stagger(const KbAugGeom& g, int B, int* groups, int* stagger) {
  static const char* g_env = getenv("KB_CHOL_GROUPS");          // debug aids
  static const char* s_env = getenv("KB_CHOL_STAGGER");
  int ng = 1, st = 0;
  if (g.nb >= 8 && B >= 32) { ng = KB_CHOL_GROUPS_DEFAULT; st = g.nb / (2 * ng) > 0 ? g.nb / (2 * ng) : 1; }
  if (g_env && atoi(g_env) > 0) ng = atoi(g_env);
  if (s_env && atoi(s_env) >= 0) st = atoi(s_env);
  if (ng > B) ng = B;
  if (ng <= 1 || st <= 0) { ng = 1; st = 0; }
  *groups = ng;
  *stagger = st;
}

void kb_chol_build_tasks(const KbAugGeom& g, int B, int num_ctas, std::vector<int4>& out) {
  out.clear();
  struct T3 { int i, k, type; };
  int ngroups = 1, stagger = 0;
  kb_chol_stagger(g, B, &ngroups, &stagger);
  // group j = candidates [B j / ngroups, B (j + 1) / ngroups): the chain workers' home candidates (b = worker + j *
  // workers) then fall into different groups
  int b_lo = 0, b_hi = B;
  auto emit = [&](const std::vector<T3>& seg) {
    for (int b = b_lo; b < b_hi; ++b)
      for (const T3& t : seg) out.push_back(make_int4(b, t.i, t.k, t.type));
  };
  const int rhs0 = g.nb, id0 = g.nb + g.sb;
  const bool thin_rhs = (g.sb == 1 && g.q <= 16);
  // the last tile row of R as short tasks when at most 56 of its rows are live (sites + seated right-hand sides)
  static const char* short_env = getenv("KB_CHOL_SHORT");       // debug aid: 0 = off
  const int live_last = g.n - (g.nb - 1) * KB_TILE + (g.seated ? g.q : 0);
  const bool short_last = live_last <= 56 && !(short_env && atoi(short_env) == 0);
  const int last_type = short_last ? KB_TASK_OFFS : KB_TASK_OFF;
  // tile column k of the current group (k == nb: the Schur tiles behind the last column)
  auto emit_column = [&](int k) {
    if (k == g.nb) {
      std::vector<T3> seg;
      for (int kk = rhs0; kk < id0; ++kk)
        for (int i = kk; i < id0; ++i)
          seg.push_back({i, kk, thin_rhs ? (KB_TASK_SCHUR_THIN | (g.nb >= 3 ? (g.nb - 1) << 8 : 0)) : KB_TASK_SCHUR});
      emit(seg);
      return;
    }
    if (k + 2 < g.nb) emit({{k + 2, k, (k + 2 == g.nb - 1) ? last_type : KB_TASK_OFF}});
    std::vector<T3> seg;
    // remaining tile rows of R and the full right-hand-side rows, two tile rows per task
    {
      std::vector<int> rows;
      for (int i = k + 3; i < g.nb - (short_last ? 1 : 0); ++i) rows.push_back(i);
      if (short_last && k + 3 <= g.nb - 1) seg.push_back({g.nb - 1, k, KB_TASK_OFFS});
      if (!thin_rhs)
        for (int i = rhs0; i < id0; ++i) rows.push_back(i);
      size_t r = 0;
      // two tile rows per task only while that still leaves every CTA a task in this column
      const bool pair = (size_t)B * rows.size() / 2 >= (size_t)num_ctas;
      for (; pair && r + 1 < rows.size(); r += 2) {
        if (rows[r + 1] == rows[r] + 1) seg.push_back({rows[r], k, KB_TASK_OFF2});
        else { seg.push_back({rows[r], k, KB_TASK_OFF}); seg.push_back({rows[r + 1], k, KB_TASK_OFF}); }
      }
      for (; r < rows.size(); ++r) seg.push_back({rows[r], k, KB_TASK_OFF});
    }
    if (thin_rhs) seg.push_back({rhs0, k, KB_TASK_THIN});
    // Schur sum over all but the last tile column, so that only two chunks are left for the tail
    if (thin_rhs && g.nb >= 3 && k == g.nb - 2) seg.push_back({rhs0, rhs0, KB_TASK_SCHUR_THIN_A});
    // identity rows (fit mode): row j of I is zero left of tile column j
    for (int j = 0; j <= k && j < g.ib; ++j) seg.push_back({id0 + j, k, KB_TASK_OFF | (j << 8)});
    emit(seg);
    // early part of the PRE task of tile row k+3 (tile columns 0..k): off the critical path
    if (k + 3 < g.nb) emit({{k + 3, k, KB_TASK_PRE_A}});
    // PRE of tile row r = k+2, last tile column (r-2 = k).  Placed at the END of the column: by the
    // time the pool gets here its inputs (PRE_A of the previous column, the chain's L[k+1,k]) are
    // in, so it never parks a CTA, and the chain still runs a good part of a column ahead.
    if (k + 2 < g.nb) emit({{k + 2, k, KB_TASK_PRE | (k << 8)}});
  };
  for (int slot = 0; slot <= g.nb + stagger * (ngroups - 1); ++slot)
    for (int grp = 0; grp < ngroups; ++grp) {
      const int k = slot - stagger * grp;
      if (k < 0 || k > g.nb) continue;
      b_lo = (int)((long long)B * grp / ngroups);
      b_hi = (int)((long long)B * (grp + 1) / ngroups);
      emit_column(k);
    }
}
