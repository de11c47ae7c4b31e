// Blocked fp64 Cholesky, triangular inverse and Sigma^-1 = L^-T L^-1 on one GPU.
//
// Replaces the dense Cholesky / solve / logdet that GaussianDistribution.log_prob reaches through
// CoLA -> jnp.linalg.cholesky (src/objectives.py:76-78) and the reverse pass of it that
// jax.value_and_grad builds (src/trainer.py:126): K_bar needs Sigma^-1 explicitly.
//
// Recursive (cache-oblivious) formulation so that almost all flops land in large DMMA GEMMs:
//   potrf(A) : potrf(A11); A21 <- A21 L11^-T (recursive TRSM); A22 -= A21 A21^T (SYRK); potrf(A22)
//   trtri(L) : W21 = -W22 (L21 W11)                        (two triangular GEMMs)
//   lauum(W) : S_ij = sum_{k >= i} W_ki^T W_kj for every lower tile, one launch
// The 128 x 128 leaves are factorised AND inverted by one CTA in shared memory (4 x 4 blocking over
// 32 x 32 sub-blocks: warp-shuffle Cholesky on the diagonal sub-blocks, DMMA for everything else);
// the inverted diagonal blocks turn every leaf-level TRSM into a GEMM.
#include <cstdlib>
#include <functional>
#include <mutex>
#include <cuda.h>
#include "lfm_common.cuh"

#define NB LFM_NB
#define LB 32                       // sub-block edge inside a leaf
#define S_LD 132                    // == 4 (mod 16): conflict-free m8n8k4 fragment reads
#define WD_LD 36                    // == 4 (mod 16)
#define LEAF_THREADS 256
#define LEAF_SMEM ((NB * S_LD + 4 * LB * WD_LD + LB * 32) * 8)

__device__ __forceinline__ void leaf_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// One warp: acc(32x32) += A(32x32, row-major lda) * op(B); TB = 1: B stored [n][k], TB = 0: B stored [k][n].
template <int TB>
__device__ __forceinline__ void warp_gemm32(double (&acc)[4][4][2], const double* __restrict__ A, int lda,
                                            const double* __restrict__ B, int ldb, int lane) {
  const int fr = lane >> 2, fc = lane & 3;
#pragma unroll
  for (int k4 = 0; k4 < LB; k4 += 4) {
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = A[(8 * i + fr) * lda + k4 + fc];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = TB ? B[(8 * j + fr) * ldb + k4 + fc] : B[(k4 + fc) * ldb + 8 * j + fr];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) leaf_dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}
__device__ __forceinline__ void warp_zero32(double (&acc)[4][4][2]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
}
// C = alpha * acc + beta * C
__device__ __forceinline__ void warp_store32(const double (&acc)[4][4][2], double* C, int ldc, double alpha,
                                             double beta, int lane) {
  const int fr = lane >> 2, fc = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double* p = C + (8 * i + fr) * ldc + 8 * j + 2 * fc;
      if (beta != 0.0) {
        p[0] = alpha * acc[i][j][0] + beta * p[0];
        p[1] = alpha * acc[i][j][1] + beta * p[1];
      } else {
        p[0] = alpha * acc[i][j][0];
        p[1] = alpha * acc[i][j][1];
      }
    }
}

// Diagonal 32 x 32 sub-block: ONE warp factorises it and inverts it in the same instruction stream.
// Lane i keeps row i of the block and column i of Y = L^-1 in REGISTERS; the pivot loop is fully
// unrolled, so every register index is a compile-time constant.  Per pivot k the dependent chain is
// one shuffle broadcast of a_kk, a branch-free rsqrt, the column scale, one shared-memory exchange of
// column k (T[k][i] = L[i][k], zeros above the diagonal, read back as 16-byte broadcasts) and the
// update of a[k+1]; the remaining rank-1 updates of the row and the whole update of the inverse
// column, Y[i][c] -= L[i][k] W[k][c], are independent DFMAs that ptxas schedules into the latency
// bubbles of that chain (a second warp coupled through a shared-memory flag was 2x slower: the
// flag needs a MEMBAR per pivot).
#define DP_LD 32
// 1/sqrt(x), x > 0 normal: MUFU.RSQ64H seed + one third-order step (the arithmetic of CUDA's rsqrt()
// without its special-case branch, which would split the pivot loop into basic blocks).
__device__ __forceinline__ double leaf_rsqrt(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = fma(-(y0 * y0), x, 1.0);
  const double t = fma(e, 0.375, 0.5);
  return fma(t, y0 * e, y0);
}
__device__ __forceinline__ int warp_potrf_trtri32(double* __restrict__ D, int ldd, double* __restrict__ T,
                                                  double* __restrict__ Winv, int lane, long long* dbg = nullptr) {
  int fail = -1;
  if (dbg && lane == 0) dbg[0] = clock64();
  double a[LB], y[LB];
#pragma unroll
  for (int j = 0; j < LB; ++j) {
    a[j] = (j <= lane) ? D[lane * ldd + j] : 0.0;
    y[j] = (j == lane) ? 1.0 : 0.0;
  }
  // The next pivot a_{k+1,k+1} - l_{k+1,k}^2 lives in lane k+1, which holds both operands in registers: it
  // is broadcast from there, so the dependent chain per pivot is rsqrt -> scale -> one DFMA -> one
  // shuffle and never waits for the shared-memory exchange.
  double akk = __shfl_sync(0xffffffffu, a[0], 0);
  if (dbg && lane == 0) dbg[1] = clock64();
#pragma unroll
  for (int k = 0; k < LB; ++k) {
    if (dbg && lane == 0 && (k == 8 || k == 16 || k == 24)) dbg[1 + k / 8] = clock64();
    if (!(akk > 0.0) && fail < 0) fail = k;
    const double rk = leaf_rsqrt(akk);
    const double lik = (lane == k) ? akk * rk : a[k] * rk;
    double akk_next = 0.0;
    if (k + 1 < LB) akk_next = __shfl_sync(0xffffffffu, fma(-lik, lik, a[k + 1]), k + 1);
    a[k] = lik;
    T[k * DP_LD + lane] = (lane >= k) ? lik : 0.0;
    const double wk = y[k] * rk;
    y[k] = wk;
    __syncwarp();
    if (k + 1 < LB) {
      if ((k + 1) & 1) {
        const double l1 = T[k * DP_LD + k + 1];
        a[k + 1] = fma(-lik, l1, a[k + 1]);
        y[k + 1] = fma(-l1, wk, y[k + 1]);
      }
#pragma unroll
      for (int j = (k + 2) & ~1; j < LB; j += 2) {
        const double2 l2 = *reinterpret_cast<const double2*>(T + k * DP_LD + j);
        a[j] = fma(-lik, l2.x, a[j]);
        a[j + 1] = fma(-lik, l2.y, a[j + 1]);
        y[j] = fma(-l2.x, wk, y[j]);
        y[j + 1] = fma(-l2.y, wk, y[j + 1]);
      }
    }
    akk = akk_next;
  }
  if (dbg && lane == 0) dbg[5] = clock64();
#pragma unroll
  for (int j = 0; j < LB; ++j) {
    D[lane * ldd + j] = (j <= lane) ? a[j] : 0.0;
    Winv[j * WD_LD + lane] = y[j];
  }
  if (dbg && lane == 0) dbg[6] = clock64();
  return fail;
}

// ---- 8-row slices (one m8 fragment row): the chain-critical panel / diagonal-update tiles are split over the
// four schedulers of the SM, 8 rows per warp, so that they cost a quarter of a 32 x 32 x 32 warp product.
template <int TB>
__device__ __forceinline__ void warp_gemm8(double (&acc)[4][2], const double* __restrict__ A, int lda,
                                           const double* __restrict__ B, int ldb, int lane) {
  const int fr = lane >> 2, fc = lane & 3;
#pragma unroll
  for (int k4 = 0; k4 < LB; k4 += 4) {
    const double a = A[fr * lda + k4 + fc];
    double b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = TB ? B[(8 * j + fr) * ldb + k4 + fc] : B[(k4 + fc) * ldb + 8 * j + fr];
#pragma unroll
    for (int j = 0; j < 4; ++j) leaf_dmma(acc[j][0], acc[j][1], a, b[j]);
  }
}
__device__ __forceinline__ void warp_store8(const double (&acc)[4][2], double* C, int ldc, double alpha, double beta,
                                            int lane) {
  const int fr = lane >> 2, fc = lane & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double* p = C + fr * ldc + 8 * j + 2 * fc;
    if (beta != 0.0) {
      p[0] = alpha * acc[j][0] + beta * p[0];
      p[1] = alpha * acc[j][1] + beta * p[1];
    } else {
      p[0] = alpha * acc[j][0];
      p[1] = alpha * acc[j][1];
    }
  }
}
__device__ __forceinline__ void warp_load32(double (&acc)[4][4][2], const double* C, int ldc, int lane) {
  const int fr = lane >> 2, fc = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double* p = C + (8 * i + fr) * ldc + 8 * j + 2 * fc;
      acc[i][j][0] = p[0];
      acc[i][j][1] = p[1];
    }
}

__device__ __forceinline__ void leaf_cp16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
// warps 1,2,3,5,6,7 (the "side" warps) synchronise among themselves inside a window
__device__ __forceinline__ void leaf_side_barrier() { asm volatile("bar.sync 1, 192;" ::: "memory"); }

// Leaf: in-place Cholesky of one 128 x 128 diagonal block AND its inverse, one CTA of 8 warps.
// Blocked 4 x 4 over 32 x 32 sub-blocks (i, j).  The dependent chain is
//   D(k) [warp 0: factor + invert the diagonal sub-block]  ->  P(k+1, k) = A(k+1,k) Winv_kk^T  ->
//   U(k+1, k+1) -= L(k+1,k) L(k+1,k)^T  ->  D(k+1),
// with P and U split into 8-row slices over warps 0-3 (one per scheduler).  Everything else -- the
// other panel tiles, the other trailing updates, the off-diagonal blocks of the inverse
// (W_ij = -Winv_ii sum_{k=j}^{i-1} L_ik W_kj, accumulated in the free upper block (j, i)) and the
// stores of finished block rows -- runs on warps 1-7 underneath the next D(k) (the "window"), so the
// chain is 4 D + 3 (P + U) slices + the last row of the inverse.
#define SBLK(i, j) (S + ((i)*LB) * S_LD + (j)*LB)
#define WDI(i) (Wd + (i)*LB * WD_LD)
__global__ void __launch_bounds__(LEAF_THREADS, 1) lfm_potrf_leaf_kernel(double* __restrict__ A, int64_t lda,
                                                                       double* __restrict__ W, int64_t ldw,
                                                                       int* __restrict__ info, int pivot_base,
                                                                       long long* __restrict__ stamps) {
  extern __shared__ __align__(16) double S[];   // [NB][S_LD], then Wd[4][LB][WD_LD], then T[LB][DP_LD]
  int nstamp = 0;
#define LEAF_STAMP() do { if (stamps && threadIdx.x == 0) stamps[nstamp++] = clock64(); } while (0)
  LEAF_STAMP();
  double* Wd = S + NB * S_LD;
  double* T = Wd + 4 * LB * WD_LD;
  __shared__ int failed;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) failed = -1;
  // ---- load: lower-triangle sub-blocks by cp.async; block (0,0) is its own group so that D(0) starts early
  for (int q = tid; q < LB * (LB / 2); q += LEAF_THREADS) {
    const int r = q >> 4, c = (q & 15) * 2;
    leaf_cp16(S + r * S_LD + c, A + (int64_t)r * lda + c);
  }
  asm volatile("cp.async.commit_group;\n" ::);
  for (int q = tid; q < NB * (NB / 2); q += LEAF_THREADS) {
    const int r = q >> 6, c = (q & 63) * 2;
    if (r >= LB && (c >> 5) <= (r >> 5)) leaf_cp16(S + r * S_LD + c, A + (int64_t)r * lda + c);
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 1;\n" ::);
  __syncthreads();
  LEAF_STAMP();

  // finished block row `i` of L (zeros right of the diagonal block) -> global; one warp
  auto store_L_row = [&](int i) {
    for (int r = 0; r < LB; ++r) {
      const int row = i * LB + r;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = h * 64 + lane * 2;
        double2 v = make_double2(0.0, 0.0);
        if ((c >> 5) <= i) v = *reinterpret_cast<const double2*>(S + row * S_LD + c);
        *reinterpret_cast<double2*>(A + (int64_t)row * lda + c) = v;
      }
    }
  };
  // finished block row `i` of W = L^-1 -> global; one warp
  auto store_W_row = [&](int i) {
    for (int r = 0; r < LB; ++r) {
      const int row = i * LB + r;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = h * 64 + lane * 2;
        const int bj = c >> 5;
        double2 v = make_double2(0.0, 0.0);
        if (bj == i) v = *reinterpret_cast<const double2*>(WDI(i) + r * WD_LD + (c & 31));
        else if (bj < i) v = *reinterpret_cast<const double2*>(S + (bj * LB + r) * S_LD + i * LB + (c & 31));
        *reinterpret_cast<double2*>(W + (int64_t)row * ldw + c) = v;
      }
    }
  };
  auto op_P = [&](int i, int k) {  // L_ik = A_ik Winv_kk^T, in place
    double acc[4][4][2];
    warp_zero32(acc);
    warp_gemm32<1>(acc, SBLK(i, k), S_LD, WDI(k), WD_LD, lane);
    __syncwarp();
    warp_store32(acc, SBLK(i, k), S_LD, 1.0, 0.0, lane);
  };
  auto op_U = [&](int i, int j, int k) {  // A_ij -= L_ik L_jk^T
    double acc[4][4][2];
    warp_zero32(acc);
    warp_gemm32<1>(acc, SBLK(i, k), S_LD, SBLK(j, k), S_LD, lane);
    warp_store32(acc, SBLK(i, j), S_LD, -1.0, 1.0, lane);
  };
  // T_ij (+)= sum_{k=k0}^{k1} L_ik W_kj into the upper block (j, i); W_jj lives in Wd, W_kj (k > j) in block (j, k)
  auto op_T = [&](int i, int j, int k0, int k1, bool accumulate) {
    double acc[4][4][2];
    if (accumulate) warp_load32(acc, SBLK(j, i), S_LD, lane);
    else warp_zero32(acc);
    for (int k = k0; k <= k1; ++k) {
      if (k == j) warp_gemm32<0>(acc, SBLK(i, k), S_LD, WDI(j), WD_LD, lane);
      else warp_gemm32<0>(acc, SBLK(i, k), S_LD, SBLK(j, k), S_LD, lane);
    }
    __syncwarp();
    warp_store32(acc, SBLK(j, i), S_LD, 1.0, 0.0, lane);
  };
  auto op_F = [&](int i, int j) {  // W_ij = -Winv_ii T_ij, in place in block (j, i)
    double acc[4][4][2];
    warp_zero32(acc);
    warp_gemm32<0>(acc, WDI(i), WD_LD, SBLK(j, i), S_LD, lane);
    __syncwarp();
    warp_store32(acc, SBLK(j, i), S_LD, -1.0, 0.0, lane);
  };

#pragma unroll 1
  for (int kb = 0; kb < 4; ++kb) {
    // ---- window kb: warp 0 runs D(kb); the other warps run the deferred work of step kb-1 ------------------
    if (warp == 0) {
      const int f = warp_potrf_trtri32(SBLK(kb, kb), S_LD, T, WDI(kb), lane, (stamps && kb == 0) ? stamps + 16 : nullptr);
      if (lane == 0 && f >= 0 && failed < 0) failed = kb * LB + f;
      if (kb == 0) asm volatile("cp.async.wait_group 0;\n" ::);
    } else if (warp == 4) {
      // shares scheduler 0 with the chain warp: stays idle
      if (kb == 0) asm volatile("cp.async.wait_group 0;\n" ::);
    } else {
      if (kb == 0) {
        asm volatile("cp.async.wait_group 0;\n" ::);
      } else if (kb == 1) {
        // P(2,0), P(3,0), T(1,0) first; then the five remaining updates of step 0
        if (warp == 1) op_P(2, 0);
        else if (warp == 2) op_P(3, 0);
        else if (warp == 3) op_T(1, 0, 0, 0, false);
        leaf_side_barrier();
        if (warp == 1) op_U(2, 1, 0);
        else if (warp == 2) op_U(3, 1, 0);
        else if (warp == 3) op_U(2, 2, 0);
        else if (warp == 5) op_U(3, 2, 0);
        else if (warp == 6) op_U(3, 3, 0);
        else if (warp == 7) op_T(2, 0, 0, 0, false);   // L20 W00 (P(2,0) is done)
        if (warp == 7) { store_L_row(0); store_W_row(0); }
      } else if (kb == 2) {
        if (warp == 1) op_P(3, 1);
        else if (warp == 2) op_F(1, 0);
        else if (warp == 3) op_T(2, 1, 1, 1, false);
        else if (warp == 5) op_T(3, 0, 0, 0, false);   // L30 W00
        leaf_side_barrier();
        if (warp == 1) op_U(3, 2, 1);
        else if (warp == 2) op_U(3, 3, 1);
        else if (warp == 3) op_T(2, 0, 1, 1, true);    // + L21 W10
        else if (warp == 5) op_T(3, 1, 1, 1, false);   // L31 W11
        else if (warp == 6) op_T(3, 0, 1, 1, true);    // + L31 W10
        else if (warp == 7) store_L_row(1);
      } else {
        if (warp == 1) op_F(2, 1);
        else if (warp == 2) op_F(2, 0);
        else if (warp == 3) op_T(3, 2, 2, 2, false);   // L32 W22
        leaf_side_barrier();
        if (warp == 1) op_T(3, 1, 2, 2, true);         // + L32 W21
        else if (warp == 2) op_T(3, 0, 2, 2, true);    // + L32 W20
        else if (warp == 5) store_W_row(2);
        else if (warp == 6) store_L_row(2);
        else if (warp == 7) store_W_row(1);            // W row 1 was finished in window 2
      }
    }
    __syncthreads();
    LEAF_STAMP();
    if (kb == 3) break;
    // ---- chain-critical slices: P(kb+1, kb) then U(kb+1, kb+1), 8 rows per warp on warps 0-3 -------------
    if (warp < 4) {
      double* C = SBLK(kb + 1, kb) + (8 * warp) * S_LD;
      double acc[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
      warp_gemm8<1>(acc, C, S_LD, WDI(kb), WD_LD, lane);
      __syncwarp();
      warp_store8(acc, C, S_LD, 1.0, 0.0, lane);
    }
    __syncthreads();
    if (warp < 4) {
      double acc[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
      warp_gemm8<1>(acc, SBLK(kb + 1, kb) + (8 * warp) * S_LD, S_LD, SBLK(kb + 1, kb), S_LD, lane);
      warp_store8(acc, SBLK(kb + 1, kb + 1) + (8 * warp) * S_LD, S_LD, -1.0, 1.0, lane);
    }
    __syncthreads();
    LEAF_STAMP();
  }
  // ---- tail: last block row of the inverse, W_3j = -Winv_33 T_3j, then the last block rows go out ----------
  if (warp >= 1 && warp <= 3) op_F(3, warp - 1);
  else if (warp >= 4) {
    // L row 3: 32 rows over warps 4-7
    for (int r = (warp - 4) * 8; r < (warp - 4) * 8 + 8; ++r) {
      const int row = 3 * LB + r;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = h * 64 + lane * 2;
        *reinterpret_cast<double2*>(A + (int64_t)row * lda + c) = *reinterpret_cast<const double2*>(S + row * S_LD + c);
      }
    }
  }
  __syncthreads();
  LEAF_STAMP();
  {
    // W row 3: 32 rows x 128 columns over all 8 warps (4 rows per warp)
    for (int r = warp * 4; r < warp * 4 + 4; ++r) {
      const int row = 3 * LB + r;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = h * 64 + lane * 2;
        const int bj = c >> 5;
        double2 v;
        if (bj == 3) v = *reinterpret_cast<const double2*>(WDI(3) + r * WD_LD + (c & 31));
        else v = *reinterpret_cast<const double2*>(S + (bj * LB + r) * S_LD + 3 * LB + (c & 31));
        *reinterpret_cast<double2*>(W + (int64_t)row * ldw + c) = v;
      }
    }
  }
  LEAF_STAMP();
  if (tid == 0 && failed >= 0) atomicCAS(info, 0, pivot_base + failed + 1);
}

static LfmSmemConfig g_leaf_smem;
static int leaf(cudaStream_t st, double* A, int64_t lda, double* W, int64_t ldw, int* info, int64_t pivot_base) {
  LFM_CUDA_OK(lfm_ensure_smem(lfm_potrf_leaf_kernel, g_leaf_smem, LEAF_SMEM));
  lfm_potrf_leaf_kernel<<<1, LEAF_THREADS, LEAF_SMEM, st>>>(A, lda, W, ldw, info, (int)pivot_base, nullptr);
  LFM_LAUNCHED(1);
  LFM_CUDA_OK(cudaGetLastError());
  return LFM_OK;
}

static inline int64_t split(int64_t n) { return (n / NB / 2) * NB; }

static LfmGemm mk(int ta, int tb, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B,
                  int64_t ldb, double* C, int64_t ldc, double alpha, double beta, int lower, int kmode) {
  LfmGemm g;
  g.transA = ta; g.transB = tb; g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.alpha = alpha; g.beta = beta; g.lower_only = lower; g.kmode = kmode;
  g.batch = 1; g.strideA = g.strideB = g.strideC = 0;
  return g;
}

// Low-priority products of the interleaved inverse, K-CHUNKED: one launch per `chunk` of the k-range (the first with the
// caller's beta, the rest accumulating).  A 64 x 64-tile CTA of a node product with K = 1024 lives for ~75 us; a dozen of
// such launches had every CTA slot of the bulk partition taken just when a trailing update of the chain-bound half became
// ready, and stream priority only acts when a slot frees (CUPTI timeline: the update waited 67 us, the chain behind it).
// In chunks the CTAs live ~20-40 us and the high-priority stream gets the next free slot.
static int64_t tri_kchunk() {
  static int64_t v = -1;
  if (v < 0) { const char* e = getenv("LFM_TRI_KCHUNK"); v = e ? atoll(e) : 0; }
  return v;
}
// Low-priority launches of the bulk partition (inverse products, early W11^T W11) take at most ONE CTA slot per SM: a
// 64 x 64-tile CTA (80 KB of shared memory, two per SM) is launched with 40 KB of padding, so that two of them do not fit
// on an SM but one of them and one CTA of a trailing update / panel do.  Priority only decides who gets a slot that
// frees; with both slots of every SM held by 40-150 us CTAs of the inverse, the short kernels of the chain-bound half
// waited for slots (CUPTI timeline: leaf-to-leaf periods of 60-170 us instead of 42).
static int low_pad() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LFM_LOW_PAD"); v = e ? atoi(e) : 0; }
  return v;
}
static int gemm_kchunked(cudaStream_t st, LfmGemm g, int64_t chunk) {
  if (chunk <= 0 || g.K <= chunk) return lfm_dgemm(st, g);
  for (int64_t k0 = 0; k0 < g.K; k0 += chunk) {
    g.k_lo = k0;
    g.k_hi = k0 + chunk;
    LFM_TRY(lfm_dgemm(st, g));
    g.beta = 1.0;
  }
  return LFM_OK;
}

// X L^T = B in place; B is m x n (ldb), L n x n lower with inverted diagonal blocks in Wd.
static int trsm_rec(cudaStream_t st, int64_t m, int64_t n, double* B, int64_t ldb, const double* L, int64_t ldl,
                    const double* Wd, int64_t ldw) {
  if (n == NB) {
    // B <- B W_kk^T  (single column tile: each CTA reads only the rows it overwrites)
    return lfm_dgemm(st, mk(0, 1, m, NB, NB, B, ldb, Wd, ldw, B, ldb, 1.0, 0.0, 0, LFM_K_FULL));
  }
  const int64_t n1 = split(n), n2 = n - n1;
  LFM_TRY(trsm_rec(st, m, n1, B, ldb, L, ldl, Wd, ldw));
  // B2 -= B1 L21^T
  LFM_TRY(lfm_dgemm(st, mk(0, 1, m, n2, n1, B, ldb, L + n1 * ldl, ldl, B + n1, ldb, -1.0, 1.0, 0, LFM_K_FULL)));
  return trsm_rec(st, m, n2, B + n1, ldb, L + n1 * ldl + n1, ldl, Wd + n1 * ldw + n1, ldw);
}

// ---- chain step: L(k+1,k) = A(k+1,k) W_kk^T and A(k+1,k+1) -= L(k+1,k) L(k+1,k)^T in ONE launch -----------------
// A cluster of 8 CTAs owns the 128 x 128 block row: CTA r computes rows [16 r, 16 r + 16) of the panel with all
// of K = 128 resident in shared memory (one cp.async sweep in four k-quarters, no multi-stage refill), writes
// them back in place, the cluster synchronises (release / acquire at cluster scope, so the rows every CTA wrote
// are visible through L2), and CTA r then updates its 16 rows of the next diagonal block against the whole
// panel.  Replaces two 16 x 128-tile GEMM launches on the dependent chain of the factorisation.
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define CS_LD 132
#define CS_SMEM ((16 + NB) * CS_LD * 8)
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256, 1)
    lfm_chain_step_kernel(double* __restrict__ P, int64_t ld, const double* __restrict__ Wkk, int64_t ldw,
                          double* __restrict__ Cd) {
  extern __shared__ __align__(16) double cs[];
  double* sA = cs;                 // [16][CS_LD]
  double* sB = cs + 16 * CS_LD;    // [128][CS_LD], stored [n][k]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const int r0 = 16 * rank;
  auto load_quarters = [&](const double* Bsrc, int64_t ldb, bool with_a) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (with_a) {
        // A rows: 16 x 32 doubles of this quarter = 256 chunks of 16 bytes
        const int r = tid >> 4, c = q * 32 + (tid & 15) * 2;
        leaf_cp16(sA + r * CS_LD + c, P + (int64_t)(r0 + r) * ld + c);
      }
      // B: 128 x 32 doubles of this quarter = 2048 chunks
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int id = tid + 256 * i;
        const int r = id >> 4, c = q * 32 + (id & 15) * 2;
        leaf_cp16(sB + r * CS_LD + c, Bsrc + (int64_t)r * ldb + c);
      }
      asm volatile("cp.async.commit_group;\n" ::);
    }
  };
  // one warp: rows [0,16) x columns [16 warp, 16 warp + 16); acc (+)= sgn * sA sB^T over the resident K = 128
  auto sweep = [&](double (&acc)[2][2][2], double sgn) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q == 0) asm volatile("cp.async.wait_group 3;\n" ::);
      else if (q == 1) asm volatile("cp.async.wait_group 2;\n" ::);
      else if (q == 2) asm volatile("cp.async.wait_group 1;\n" ::);
      else asm volatile("cp.async.wait_group 0;\n" ::);
      __syncthreads();
#pragma unroll
      for (int k4 = q * 32; k4 < q * 32 + 32; k4 += 4) {
        const double a0 = sgn * sA[fr * CS_LD + k4 + fc];
        const double a1 = sgn * sA[(8 + fr) * CS_LD + k4 + fc];
        const double b0 = sB[(16 * warp + fr) * CS_LD + k4 + fc];
        const double b1 = sB[(16 * warp + 8 + fr) * CS_LD + k4 + fc];
        leaf_dmma(acc[0][0][0], acc[0][0][1], a0, b0);
        leaf_dmma(acc[0][1][0], acc[0][1][1], a0, b1);
        leaf_dmma(acc[1][0][0], acc[1][0][1], a1, b0);
        leaf_dmma(acc[1][1][0], acc[1][1][1], a1, b1);
      }
    }
  };
  // ---- phase 1: panel rows
  load_quarters(Wkk, ldw, true);
  double acc[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  sweep(acc, 1.0);
  __syncthreads();  // every warp is done reading sA / sB
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = 8 * i + fr, c = 16 * warp + 8 * j + 2 * fc;
      const double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
      *reinterpret_cast<double2*>(P + (int64_t)(r0 + r) * ld + c) = v;   // L(k+1,k) rows, in place
      *reinterpret_cast<double2*>(sA + r * CS_LD + c) = v;               // and as the A operand of phase 2
    }
  __threadfence();
  cluster.sync();
  // ---- phase 2: rows of the next diagonal block (lower part only: columns <= last row of this CTA)
  load_quarters(P, ld, false);
  const bool live = 16 * warp <= r0 + 15;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = 8 * i + fr, c = 16 * warp + 8 * j + 2 * fc;
      double2 v = make_double2(0.0, 0.0);
      if (live) v = *reinterpret_cast<const double2*>(Cd + (int64_t)(r0 + r) * ld + c);
      acc[i][j][0] = v.x; acc[i][j][1] = v.y;
    }
  sweep(acc, -1.0);
  if (live) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = 8 * i + fr, c = 16 * warp + 8 * j + 2 * fc;
        *reinterpret_cast<double2*>(Cd + (int64_t)(r0 + r) * ld + c) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
  }
}
static int chain_step(cudaStream_t st, double* P, int64_t ld, const double* Wkk, int64_t ldw, double* Cd) {
  static LfmSmemConfig smem_cfg;
  LFM_CUDA_OK(lfm_ensure_smem(lfm_chain_step_kernel, smem_cfg, CS_SMEM));
  lfm_chain_step_kernel<<<8, 256, CS_SMEM, st>>>(P, ld, Wkk, ldw, Cd);
  LFM_LAUNCHED(1);
  LFM_CUDA_OK(cudaGetLastError());
  return LFM_OK;
}
static int chain_fused_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LFM_CHAIN_FUSED"); v = e ? atoi(e) : 1; }
  return v;
}

// Right-looking blocked Cholesky with 128-wide panels and one step of look-ahead for small / medium n.
// The dependent chain  leaf(k) -> L(k+1,k) = A(k+1,k) W_kk^T -> A(k+1,k+1) -= L(k+1,k) L(k+1,k)^T -> leaf(k+1)
// runs on an internal stream of its own SM partition (below); the rest of the panel and of the trailing update of
// step k runs on the bulk stream underneath leaf(k+1):
//   bulk : wait leaf(k) | rows >= k+2 of the panel | wait L(k+1,k) | trailing update minus block (k+1,k+1)
//   chain: wait bulk(k-1) before the panel of step k reads A(k+1, k)
// Fork/join is by events only (no host synchronisation; legal under stream capture).  Without the partition the chain
// stream is a high-priority stream and the bulk stream is the caller's.
struct LookAhead {
  cudaStream_t chain = nullptr, tri = nullptr;
  cudaStream_t pan = nullptr;    // panel stream: rows >= k+2 of panel k, underneath the trailing update of step k-1 (highest priority)
  cudaStream_t ua = nullptr;     // early part of a trailing update (block column k+1): beside the late part, not in front of it
  cudaStream_t fill = nullptr;   // lowest priority: work that only fills idle SMs of the chain-bound half (early W11^T W11)
  cudaStream_t bulk = nullptr;   // SM partition only: the bulk stream of the large partition (else the caller's stream)
  cudaStream_t tri2 = nullptr;   // SM partition only: a SUBSET of the large partition for the one long product of the inverse
  cudaEvent_t fork = nullptr, join = nullptr, tri_join = nullptr, bulk_join = nullptr, tri_mid = nullptr, tri2_join = nullptr;
  cudaEvent_t leaf_done[2] = {nullptr, nullptr}, p1_done[2] = {nullptr, nullptr}, bulk_done[2] = {nullptr, nullptr};
  cudaEvent_t ua_done[2] = {nullptr, nullptr}, ud_done[2] = {nullptr, nullptr}, pan_done[2] = {nullptr, nullptr}, pan_join = nullptr, ua_join = nullptr, w11_done = nullptr, fill_join = nullptr;
  bool ok = false;
  int chain_sms = 0;             // SMs of the chain partition (0: no partition, priorities only)
  // SM partition (green contexts, driver API >= 12.4, reached through cudaGetDriverEntryPoint so that the library
  // has no link-time dependency on libcuda): the leaf needs a whole SM (all registers, 170 KB of shared memory) and
  // the chain step a cluster of 8 such SMs, so on a device that the 64 x 64-tile bulk GEMMs keep full -- two CTAs per
  // SM, refilled slot by slot -- a pending chain kernel only starts when a bulk launch drains, whatever its stream
  // priority (measured: the chain waited for the tail of every trailing update, and 265 us for the 2048^3 product of
  // the inverse in the middle of an N = 4096 factorisation).  The chain therefore gets 8 SMs of its own (one
  // cluster-capable group); the trailing updates and the inverse share the other 140.
  // streams of this host thread inside the process-wide SM partition of the device (sm_partition below)
  bool init_partition(int dev, int prio_lo, int prio_hi);
  bool tried = false;
  bool init() {
    if (ok || tried) return ok;
    tried = true;
    int lo = 0, hi = 0, dev = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess || cudaGetDevice(&dev) != cudaSuccess) return false;
    if (!init_partition(dev, lo, hi)) {
      bulk = nullptr; tri2 = nullptr; chain_sms = 0;
      if (cudaStreamCreateWithPriority(&chain, cudaStreamNonBlocking, hi) != cudaSuccess) return false;
      if (cudaStreamCreateWithPriority(&tri, cudaStreamNonBlocking, lo) != cudaSuccess) return false;
      if (cudaStreamCreateWithPriority(&pan, cudaStreamNonBlocking, hi) != cudaSuccess) return false;
      if (cudaStreamCreateWithPriority(&ua, cudaStreamNonBlocking, hi) != cudaSuccess) return false;
      if (cudaStreamCreateWithPriority(&fill, cudaStreamNonBlocking, lo) != cudaSuccess) return false;
    }
    cudaEvent_t* all[] = {&fork, &join, &tri_join, &bulk_join, &tri_mid, &tri2_join, &leaf_done[0], &leaf_done[1], &p1_done[0], &p1_done[1],
                          &bulk_done[0], &bulk_done[1], &ua_done[0], &ua_done[1], &ud_done[0], &ud_done[1], &pan_done[0], &pan_done[1], &pan_join, &ua_join, &w11_done, &fill_join};
    for (cudaEvent_t* e : all)
      if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return false;
    ok = true;
    return true;
  }
};
static thread_local LookAhead g_la_dev[16];  // per host thread and device

// Process-wide SM partition of one device: the green contexts are created once (first use, any thread) and live as
// long as the process; every host thread creates its own streams inside them.
struct SmPartition {
  std::once_flag once;
  bool ok = false;
  bool have_sub = false;
  int chain_sms = 0;
  CUgreenCtx chain_ctx{}, bulk_ctx{}, sub_ctx{};
  CUresult (*stream_create)(CUstream*, CUgreenCtx, unsigned, int) = nullptr;
};
static SmPartition g_sm_partition[16];

static void sm_partition_create(SmPartition& P, int dev) {
  const char* env = getenv("LFM_SM_PARTITION");
  if (env && atoi(env) == 0) return;
  // Nsight Compute cannot replay kernels launched into a green context ("Failed to prepare kernel for profiling"):
  // under ncu the chain falls back to stream priorities (ncu serialises the launches anyway)
  if (!env && (getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NV_TPS_LAUNCH_TOKEN"))) return;
  CUresult (*pDevGet)(CUdevice*, int) = nullptr;
  CUresult (*pGetRes)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*pSplit)(CUdevResource*, unsigned*, const CUdevResource*, CUdevResource*, unsigned, unsigned) = nullptr;
  CUresult (*pDesc)(CUdevResourceDesc*, CUdevResource*, unsigned) = nullptr;
  CUresult (*pCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned) = nullptr;
  CUresult (*pGreenRes)(CUgreenCtx, CUdevResource*, CUdevResourceType) = nullptr;
  auto ep = [](const char* name, void** fn) {
    cudaDriverEntryPointQueryResult q;
    return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && *fn != nullptr;
  };
  if (!ep("cuDeviceGet", (void**)&pDevGet) || !ep("cuDeviceGetDevResource", (void**)&pGetRes) ||
      !ep("cuDevSmResourceSplitByCount", (void**)&pSplit) || !ep("cuDevResourceGenerateDesc", (void**)&pDesc) ||
      !ep("cuGreenCtxCreate", (void**)&pCreate) || !ep("cuGreenCtxStreamCreate", (void**)&P.stream_create)) {
    cudaGetLastError();
    return;
  }
  CUdevice cudev;
  CUdevResource all, grp[1], rem;
  unsigned nb = 1;
  if (pDevGet(&cudev, dev) != CUDA_SUCCESS || pGetRes(cudev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return;
  if (all.sm.smCount < 64) return;
  if (pSplit(grp, &nb, &all, &rem, 0, 8) != CUDA_SUCCESS || nb != 1 || grp[0].sm.smCount < 8 || rem.sm.smCount < 32) return;
  CUdevResourceDesc dA, dB;
  if (pDesc(&dA, &grp[0], 1) != CUDA_SUCCESS || pDesc(&dB, &rem, 1) != CUDA_SUCCESS) return;
  if (pCreate(&P.chain_ctx, dA, cudev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return;
  if (pCreate(&P.bulk_ctx, dB, cudev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return;
  P.chain_sms = (int)grp[0].sm.smCount;
  P.ok = true;
  // A third context on HALF of the large partition (its SMs belong to both): the top node of the interleaved inverse
  // issues T = L21 W11, (n/2)^3 flops in one launch of long-lived 64 x 64-tile CTAs, in the middle of the
  // factorisation.  On the whole partition it took every slot for 280 us and the (short) trailing updates of the
  // chain-bound second half queued behind it; confined to half of the SMs it runs underneath that half instead.
  const char* e2 = getenv("LFM_TRI2_SMS");
  const int want = e2 ? atoi(e2) : 88;
  CUdevResource resB, sub[1], rem2;
  unsigned nb2 = 1;
  CUdevResourceDesc dC;
  if (want > 0 && ep("cuGreenCtxGetDevResource", (void**)&pGreenRes) &&
      pGreenRes(P.bulk_ctx, &resB, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS &&
      pSplit(sub, &nb2, &resB, &rem2, 0, (unsigned)want) == CUDA_SUCCESS && nb2 == 1 &&
      pDesc(&dC, &sub[0], 1) == CUDA_SUCCESS && pCreate(&P.sub_ctx, dC, cudev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS)
    P.have_sub = true;
  else
    cudaGetLastError();
}

bool LookAhead::init_partition(int dev, int prio_lo, int prio_hi) {
  if (dev < 0 || dev >= 16) return false;
  SmPartition& P = g_sm_partition[dev];
  std::call_once(P.once, sm_partition_create, std::ref(P), dev);
  if (!P.ok) return false;
  CUstream sc, sb, stri, st2, sp;
  // priorities inside the large partition: panel stream > trailing updates > inverse (numerically lower = higher)
  const int prio_bulk = (prio_hi + 1 < prio_lo) ? prio_hi + 1 : prio_hi;
  if (P.stream_create(&sc, P.chain_ctx, CU_STREAM_NON_BLOCKING, prio_hi) != CUDA_SUCCESS) return false;
  if (P.stream_create(&sb, P.bulk_ctx, CU_STREAM_NON_BLOCKING, prio_bulk) != CUDA_SUCCESS) return false;
  // the inverse sits one level above the lowest priority, which belongs to the filler stream
  const int prio_tri = (prio_lo - 1 > prio_bulk) ? prio_lo - 1 : prio_lo;
  if (P.stream_create(&stri, P.bulk_ctx, CU_STREAM_NON_BLOCKING, prio_tri) != CUDA_SUCCESS) return false;
  CUstream sfill;
  if (P.stream_create(&sfill, P.bulk_ctx, CU_STREAM_NON_BLOCKING, prio_lo) != CUDA_SUCCESS) return false;
  fill = (cudaStream_t)sfill;
  if (P.stream_create(&sp, P.bulk_ctx, CU_STREAM_NON_BLOCKING, prio_hi) != CUDA_SUCCESS) return false;
  CUstream sua;
  if (P.stream_create(&sua, P.bulk_ctx, CU_STREAM_NON_BLOCKING, prio_hi) != CUDA_SUCCESS) return false;
  chain = (cudaStream_t)sc; bulk = (cudaStream_t)sb; tri = (cudaStream_t)stri; pan = (cudaStream_t)sp; ua = (cudaStream_t)sua;
  chain_sms = P.chain_sms;
  // the one long product of the inverse: its own stream (it must not sit in front of the later node products of `tri`),
  // inside the sub-partition if there is one, else on the whole bulk partition
  if (P.stream_create(&st2, P.have_sub ? P.sub_ctx : P.bulk_ctx, CU_STREAM_NON_BLOCKING, prio_tri) == CUDA_SUCCESS) tri2 = (cudaStream_t)st2;
  return true;
}

static int lookahead_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LFM_LOOKAHEAD"); v = e ? atoi(e) : 1; }
  return v;
}

static int potrf_right_looking_serial(cudaStream_t st, int64_t n, double* A, int64_t lda, double* W, int64_t ldw,
                                      int* info, int64_t pivot_base) {
  for (int64_t k = 0; k < n; k += NB) {
    double* Akk = A + k * lda + k;
    double* Wkk = W + k * ldw + k;
    LFM_TRY(leaf(st, Akk, lda, Wkk, ldw, info, pivot_base + k));
    const int64_t m = n - k - NB;
    if (m <= 0) break;
    double* P = Akk + NB * lda;  // panel below the diagonal block, m x 128
    LFM_TRY(lfm_dgemm(st, mk(0, 1, m, NB, NB, P, lda, Wkk, ldw, P, lda, 1.0, 0.0, 0, LFM_K_FULL)));
    LFM_TRY(lfm_dgemm(st, mk(0, 1, m, m, NB, P, lda, P, lda, P + NB, lda, -1.0, 1.0, 1, LFM_K_FULL)));
  }
  return LFM_OK;
}

// With `with_trtri` (n / 128 a power of two, lda == ldw) the triangular inverse W = L^-1 is built DURING the
// factorisation on a third, low-priority stream: the node of the trtri recursion tree that covers blocks
// [o, o + 2 m) needs  T = W11^T-product with L21  as soon as its left half is inverted and the panels of its
// columns are final, and  W21 = -W22 T  as soon as its right half is inverted; in the second half of the
// factorisation, where the dependent chain leaves most SMs idle, these products are free.
__global__ void lfm_chol_diag_copy_kernel(int64_t n, const double* __restrict__ A, int64_t lda, double* __restrict__ d) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) d[i] = A[i * lda + i];
}
static int early_lauum_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LFM_EARLY_LAUUM"); v = e ? atoi(e) : 2; }
  return v;
}
// `ldiag` (with_trtri only, may be NULL): n doubles that receive diag(L).  When it is given and the sweep has at least 16
// blocks, the first half of Sigma^-1 = W^T W is started EARLY: diag(L11) is copied out and  S11' = W11^T W11  overwrites
// the lower triangle of the (dead) L11 block on the lowest-priority stream of the bulk partition.  *early_done = 1 then,
// and the caller finishes with lfm_lauum_late (S11 = S11' + W21^T W21, the other blocks as usual) instead of lfm_lauum.
// LFM_EARLY_LAUUM = 2 (default): S11' starts when the chain is through, beside the serial tail of the inverse (four
// short dependent launches on a mostly idle device): N = 4000 evaluation 3.238 -> 3.222 ms.  1: as soon as W11 is complete,
// half-way through the sweep -- measured SLOWER (3.31 vs 3.25 ms): the 64 x 64-tile CTAs of the filler hold slots of the
// partition for 40-170 us and the short kernels the chain-bound half waits for queue behind them (priority only
// decides who gets a slot that frees).  0: off.
// `chain_ready` (may be NULL): an event the caller recorded on `st` once the first TWO block columns of A were complete
// while it went on filling the rest (nlml.cu builds Sigma that way): the chain -- leaf(0), chain step 0 -- starts behind
// that event, everything else behind all of st's work as usual.
static int potrf_right_looking(cudaStream_t st, int64_t n, double* A, int64_t lda, double* W, int64_t ldw, int* info,
                               int64_t pivot_base, bool with_trtri = false, double* ldiag = nullptr, int* early_done = nullptr,
                               cudaEvent_t chain_ready = nullptr) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) dev = -1;
  if (!lookahead_mode() || n < 4 * NB || dev < 0 || !g_la_dev[dev].init()) {
    LFM_TRY(potrf_right_looking_serial(st, n, A, lda, W, ldw, info, pivot_base));
    LFM_TRY(with_trtri ? lfm_trtri(st, n, A, lda, W, ldw) : LFM_OK);
    if (ldiag) {
      lfm_chol_diag_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, A, lda, ldiag);
      LFM_LAUNCHED(1);
      LFM_CUDA_OK(cudaGetLastError());
    }
    return LFM_OK;
  }
  LookAhead& la = g_la_dev[dev];
  cudaStream_t ch = la.chain;
  cudaStream_t tr = la.tri;
  cudaStream_t bk = la.bulk ? la.bulk : st;   // bulk work: the large SM partition, or the caller's stream
  LFM_CUDA_OK(cudaEventRecord(la.fork, st));
  LFM_CUDA_OK(cudaStreamWaitEvent(ch, chain_ready ? chain_ready : la.fork, 0));
  if (bk != st) LFM_CUDA_OK(cudaStreamWaitEvent(bk, la.fork, 0));
  if (with_trtri) LFM_CUDA_OK(cudaStreamWaitEvent(tr, la.fork, 0));
  const int64_t nb = n / NB;
  // one product of the trtri node with half size mb blocks whose first block is ob (see trtri_levels)
  bool tri2_used = false;
  auto tri_g1 = [&](int64_t ob, int64_t mb) -> int {  // T^T = W11^T L21^T into the strictly upper block of W
    const int64_t o = ob * NB, m = mb * NB;
    cudaStream_t s1 = tr;
    if (la.tri2 && 2 * mb == nb && nb >= 16) {   // the top node: everything the low-priority stream was told to wait for
      LFM_CUDA_OK(cudaEventRecord(la.tri_mid, tr));   // and everything it has been given so far precedes it
      LFM_CUDA_OK(cudaStreamWaitEvent(la.tri2, la.tri_mid, 0));
      s1 = la.tri2;
      tri2_used = true;
    }
    LfmGemm g = mk(1, 1, m, m, m, W + o * ldw + o, ldw, A + (o + m) * lda + o, lda, W + o * ldw + o + m, ldw, 1.0, 0.0, 0,
                   LFM_K_GE_ROW);
    g.tile = 3; g.smem_pad = low_pad();
    return gemm_kchunked(s1, g, tri_kchunk());
  };
  auto tri_g2 = [&](cudaStream_t s2, int64_t ob, int64_t mb) -> int {  // W21 = -W22 T
    const int64_t o = ob * NB, m = mb * NB;
    // (the serial tail after the join runs on the whole device with nothing to yield to: one launch)
    LfmGemm g = mk(0, 1, m, m, m, W + (o + m) * ldw + o + m, ldw, W + o * ldw + o + m, ldw, W + (o + m) * ldw + o, ldw, -1.0,
                   0.0, 0, LFM_K_LE_ROW);
    if (s2 != st) { g.tile = 3; g.smem_pad = low_pad(); }
    return gemm_kchunked(s2, g, s2 == st ? 0 : tri_kchunk());
  };
  const bool early = with_trtri && ldiag != nullptr && early_done != nullptr && nb >= 16 && nb % 2 == 0 && la.fill != nullptr &&
                     early_lauum_mode() != 0;
  if (early) LFM_CUDA_OK(cudaStreamWaitEvent(la.fill, la.fork, 0));
  int step = 0;
  bool ua_used = false, ub_used = false;
  cudaStream_t pn = la.pan, ua_s = la.ua;
  LFM_CUDA_OK(cudaStreamWaitEvent(pn, la.fork, 0));
  LFM_CUDA_OK(cudaStreamWaitEvent(ua_s, la.fork, 0));
  for (int64_t k = 0; k < n; k += NB, ++step) {
    const int e = step & 1;
    double* Akk = A + k * lda + k;
    double* Wkk = W + k * ldw + k;
    LFM_TRY(leaf(ch, Akk, lda, Wkk, ldw, info, pivot_base + k));
    const int64_t m = n - k - NB;
    LFM_CUDA_OK(cudaEventRecord(la.leaf_done[e], ch));
    if (with_trtri) {
      // nodes that END at block `step`: their right half is now inverted (lower levels first)
      // (the nodes that end at the LAST block are the serial tail of the inverse -- every level waits for the one
      // below: they run after the join, on the caller's stream, i.e. on all SMs of the device)
      LFM_CUDA_OK(cudaStreamWaitEvent(tr, la.leaf_done[e], 0));
      for (int64_t mb = 1; 2 * mb <= nb && m > 0; mb *= 2)
        if ((step + 1) % (2 * mb) == 0) LFM_TRY(tri_g2(tr, step + 1 - 2 * mb, mb));
      if (early && 2 * (step + 1) == nb) LFM_CUDA_OK(cudaEventRecord(la.w11_done, tr));
      if (early && early_lauum_mode() == 1 && 2 * (step + 1) == nb) {
        // W11 is complete behind everything `tr` has been given so far; every product that reads L11 precedes it there
        const int64_t h = n / 2;
        LFM_CUDA_OK(cudaStreamWaitEvent(la.fill, la.w11_done, 0));
        lfm_chol_diag_copy_kernel<<<(unsigned)((h + 255) / 256), 256, 0, la.fill>>>(h, A, lda, ldiag);
        LFM_LAUNCHED(1);
        LFM_CUDA_OK(cudaGetLastError());
        LfmGemm g = mk(1, 0, h, h, h, W, ldw, W, ldw, A, lda, 1.0, 0.0, 1, LFM_K_GE_ROWCOL);
        g.tile = 3;   // 64 x 64 tiles: they share an SM with the trailing updates' CTAs (a 128 x 128 tile takes a whole SM)
        g.smem_pad = low_pad();
        LFM_TRY(gemm_kchunked(la.fill, g, 512));
      }
    }
    if (m <= 0) break;
    double* P = Akk + NB * lda;  // panel below the diagonal block, m x 128
    // ---- chain: first 128 rows of the panel, then the next diagonal block.  Block column k below the diagonal block is
    // final once the EARLY part U_a of the previous trailing update is (below), not the whole update.
    if (ua_used) {
      LFM_CUDA_OK(cudaStreamWaitEvent(ch, la.ua_done[e ^ 1], 0));
      LFM_CUDA_OK(cudaStreamWaitEvent(ch, la.ud_done[e ^ 1], 0));
    }
    if (chain_fused_mode()) {
      LFM_TRY(chain_step(ch, P, lda, Wkk, ldw, P + NB));
    } else {
      LfmGemm g = mk(0, 1, NB, NB, NB, P, lda, Wkk, ldw, P, lda, 1.0, 0.0, 0, LFM_K_FULL);
      g.tile = 1;
      LFM_TRY(lfm_dgemm(ch, g));
      LfmGemm u = mk(0, 1, NB, NB, NB, P, lda, P, lda, P + NB, lda, -1.0, 1.0, 0, LFM_K_FULL);
      u.tile = 1;
      LFM_TRY(lfm_dgemm(ch, u));
    }
    LFM_CUDA_OK(cudaEventRecord(la.p1_done[e], ch));
    // nodes whose LEFT half ends at block `step`: W11 is complete and, once this step's panel is final, so is L21
    auto tri_front = [&](bool wait_panel) -> int {
      if (!with_trtri) return LFM_OK;
      LFM_CUDA_OK(cudaStreamWaitEvent(tr, la.p1_done[e], 0));
      if (wait_panel) LFM_CUDA_OK(cudaStreamWaitEvent(tr, la.pan_done[e], 0));
      for (int64_t mb = 1; 2 * mb <= nb; mb *= 2)
        if ((step + 1) % (2 * mb) == mb) LFM_TRY(tri_g1(step + 1 - mb, mb));
      return LFM_OK;
    };
    if (m <= NB) { LFM_TRY(tri_front(false)); continue; }
    // ---- panel stream: rows >= k+2 of the panel, concurrently with the chain step AND with the late part U_b of the
    // previous trailing update (highest priority in the partition: its CTAs take the slots U_b's CTAs free)
    double* P2 = P + NB * lda;  // rows >= k+2 of the panel, (m - 128) x 128
    const int64_t m2 = m - NB;
    LFM_CUDA_OK(cudaStreamWaitEvent(pn, la.leaf_done[e], 0));
    if (ua_used) LFM_CUDA_OK(cudaStreamWaitEvent(pn, la.ua_done[e ^ 1], 0));
    LFM_TRY(lfm_dgemm(pn, mk(0, 1, m2, NB, NB, P2, lda, Wkk, ldw, P2, lda, 1.0, 0.0, 0, LFM_K_FULL)));
    LFM_CUDA_OK(cudaEventRecord(la.pan_done[e], pn));
    // ---- the trailing update A(k+1:, k+1:) -= L(k+1:, k) L(k+1:, k)^T in four parts on four streams.  The chain has
    // done block (k+1, k+1).  The EARLY parts -- U_a: block column k+1 below it (stream ua), U_d: block (k+2, k+2)
    // (panel stream) -- are everything step k+1 needs for its panel, its chain step and, after that, its leaf; the
    // chain and the panel stream wait for THEM.  The LATE part U_b -- the rest, one trapezoid launch on the bulk stream
    // -- starts at the same moment and runs underneath chain step k+1, panel k+1 and leaf(k+2); the early parts of
    // step k+1 wait for it (they accumulate into blocks it writes).  While the updates are the longer side, the bulk
    // stream therefore runs U_b(k-1), U_b(k), ... back to back (before: chain step + panel sat BETWEEN two updates,
    // ~15 us of an idle partition per step for 17 steps of an N = 4096 sweep).
    const bool have_ub = m2 > NB;
    LFM_CUDA_OK(cudaStreamWaitEvent(ua_s, la.p1_done[e], 0));
    LFM_CUDA_OK(cudaStreamWaitEvent(ua_s, la.pan_done[e], 0));
    if (ub_used) LFM_CUDA_OK(cudaStreamWaitEvent(ua_s, la.bulk_done[e ^ 1], 0));
    LFM_TRY(lfm_dgemm(ua_s, mk(0, 1, m2, NB, NB, P2, lda, P, lda, P2 + NB, lda, -1.0, 1.0, 0, LFM_K_FULL)));
    LFM_CUDA_OK(cudaEventRecord(la.ua_done[e], ua_s));
    LFM_CUDA_OK(cudaStreamWaitEvent(pn, la.p1_done[e], 0));
    if (ub_used) LFM_CUDA_OK(cudaStreamWaitEvent(pn, la.bulk_done[e ^ 1], 0));
    LFM_TRY(lfm_dgemm(pn, mk(0, 1, NB, NB, NB, P2, lda, P2, lda, P2 + 2 * NB, lda, -1.0, 1.0, 1, LFM_K_FULL)));
    LFM_CUDA_OK(cudaEventRecord(la.ud_done[e], pn));
    ua_used = true;
    if (have_ub) {
      LFM_CUDA_OK(cudaStreamWaitEvent(bk, la.p1_done[e], 0));
      LFM_CUDA_OK(cudaStreamWaitEvent(bk, la.pan_done[e], 0));
      LfmGemm u = mk(0, 1, m2, m2, NB, P2, lda, P2, lda, P2 + 2 * NB, lda, -1.0, 1.0, 1, LFM_K_FULL);
      u.tri_skip = NB;   // block (k+2, k+2) is U_d's
      LFM_TRY(lfm_dgemm(bk, u));
      LFM_CUDA_OK(cudaEventRecord(la.bulk_done[e], bk));
    }
    // (bulk_done[e] of a step without a late part is never waited for: every later step is without one too)
    ub_used = have_ub;
    LFM_TRY(tri_front(true));
  }
  // join: everything after the factorisation is ordered behind the chain (and the inverse)
  LFM_CUDA_OK(cudaEventRecord(la.join, ch));
  LFM_CUDA_OK(cudaStreamWaitEvent(st, la.join, 0));
  if (early && early_lauum_mode() == 2) {
    // mode 2: S11' starts when the chain is through -- beside the serial tail of the inverse, whose first levels are
    // four short dependent launches that leave most of the device idle
    const int64_t h = n / 2;
    LFM_CUDA_OK(cudaStreamWaitEvent(la.fill, la.w11_done, 0));
    LFM_CUDA_OK(cudaStreamWaitEvent(la.fill, la.join, 0));
    lfm_chol_diag_copy_kernel<<<(unsigned)((h + 255) / 256), 256, 0, la.fill>>>(h, A, lda, ldiag);
    LFM_LAUNCHED(1);
    LFM_CUDA_OK(cudaGetLastError());
    LfmGemm g = mk(1, 0, h, h, h, W, ldw, W, ldw, A, lda, 1.0, 0.0, 1, LFM_K_GE_ROWCOL);
    g.tile = 3;
    LFM_TRY(gemm_kchunked(la.fill, g, 512));
  }
  if (bk != st) {
    LFM_CUDA_OK(cudaEventRecord(la.bulk_join, bk));
    LFM_CUDA_OK(cudaStreamWaitEvent(st, la.bulk_join, 0));
  }
  LFM_CUDA_OK(cudaEventRecord(la.pan_join, pn));
  LFM_CUDA_OK(cudaStreamWaitEvent(st, la.pan_join, 0));
  LFM_CUDA_OK(cudaEventRecord(la.ua_join, ua_s));
  LFM_CUDA_OK(cudaStreamWaitEvent(st, la.ua_join, 0));
  if (with_trtri) {
    LFM_CUDA_OK(cudaEventRecord(la.tri_join, tr));
    LFM_CUDA_OK(cudaStreamWaitEvent(st, la.tri_join, 0));
    if (tri2_used) {
      LFM_CUDA_OK(cudaEventRecord(la.tri2_join, la.tri2));
      LFM_CUDA_OK(cudaStreamWaitEvent(st, la.tri2_join, 0));
    }
    for (int64_t mb = 1; 2 * mb <= nb; mb *= 2) LFM_TRY(tri_g2(st, nb - 2 * mb, mb));
  }
  if (early) {
    LFM_CUDA_OK(cudaEventRecord(la.fill_join, la.fill));
    LFM_CUDA_OK(cudaStreamWaitEvent(st, la.fill_join, 0));
    const int64_t h = n / 2;   // diag(L22); diag(L11) went out before S11' overwrote it
    lfm_chol_diag_copy_kernel<<<(unsigned)((h + 255) / 256), 256, 0, st>>>(h, A + h * lda + h, lda, ldiag + h);
    LFM_LAUNCHED(1);
    LFM_CUDA_OK(cudaGetLastError());
    *early_done = 1;
  } else if (ldiag) {
    lfm_chol_diag_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, A, lda, ldiag);
    LFM_LAUNCHED(1);
    LFM_CUDA_OK(cudaGetLastError());
  }
  return LFM_OK;
}

static int64_t rl_threshold() {
  static int64_t v = -1;
  if (v < 0) {
    const char* e = getenv("LFM_RL_THRESHOLD");
    v = e ? atoll(e) : 4096;
  }
  return v;
}

static int potrf_rec(cudaStream_t st, int64_t n, double* A, int64_t lda, double* W, int64_t ldw, int* info,
                     int64_t pivot_base) {
  if (n == NB) return leaf(st, A, lda, W, ldw, info, pivot_base);
  if (n <= rl_threshold()) return potrf_right_looking(st, n, A, lda, W, ldw, info, pivot_base);
  const int64_t n1 = split(n), n2 = n - n1;
  LFM_TRY(potrf_rec(st, n1, A, lda, W, ldw, info, pivot_base));
  double* A21 = A + n1 * lda;
  double* A22 = A21 + n1;
  LFM_TRY(trsm_rec(st, n2, n1, A21, lda, A, lda, W, ldw));
  LFM_TRY(lfm_dgemm(st, mk(0, 1, n2, n2, n1, A21, lda, A21, lda, A22, lda, -1.0, 1.0, 1, LFM_K_FULL)));
  return potrf_rec(st, n2, A22, lda, W + n1 * ldw + n1, ldw, info, pivot_base + n1);
}

// Cholesky and W = L^-1 together: for a single right-looking sweep (n <= threshold, power-of-two block count)
// the inverse is interleaved with the factorisation, otherwise the two run back to back.
// Whether lfm_potrf_trtri_diag(n, lda == ldw) runs as ONE interleaved right-looking sweep (the path that can start its chain
// behind a `chain_ready` event).
bool lfm_potrf_trtri_is_one_sweep(int64_t n) {
  const int64_t nblk = n / NB;
  static int fuse = -1;
  if (fuse < 0) { const char* e = getenv("LFM_FUSE_TRTRI"); fuse = e ? atoi(e) : 1; }
  return fuse && n > 0 && n % NB == 0 && nblk >= 4 && (nblk & (nblk - 1)) == 0 && n <= rl_threshold() && lookahead_mode();
}
// `chain_ready` (may be NULL, only with lfm_potrf_trtri_is_one_sweep(n)): see potrf_right_looking; the caller has then zeroed
// *info itself BEFORE recording the event (the first leaf may report a failing pivot as soon as the event fires).
int lfm_potrf_trtri_diag(cudaStream_t st, int64_t n, double* A, int64_t lda, double* W, int64_t ldw, int* info,
                         double* ldiag, int* early_done, cudaEvent_t chain_ready) {
  if (n <= 0 || n % NB) return LFM_ERR_INVALID;
  if (early_done) *early_done = 0;
  if (lfm_potrf_trtri_is_one_sweep(n) && lda == ldw) {
    if (!chain_ready) LFM_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int), st));
    return potrf_right_looking(st, n, A, lda, W, ldw, info, 0, true, ldiag, early_done, chain_ready);
  }
  if (chain_ready) return LFM_ERR_INVALID;
  LFM_TRY(lfm_potrf(st, n, A, lda, W, ldw, info));
  LFM_TRY(lfm_trtri(st, n, A, lda, W, ldw));
  if (ldiag) {
    lfm_chol_diag_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, A, lda, ldiag);
    LFM_LAUNCHED(1);
    LFM_CUDA_OK(cudaGetLastError());
  }
  return LFM_OK;
}
int lfm_potrf_trtri(cudaStream_t st, int64_t n, double* A, int64_t lda, double* W, int64_t ldw, int* info) {
  return lfm_potrf_trtri_diag(st, n, A, lda, W, ldw, info, nullptr, nullptr, nullptr);
}

int lfm_potrf(cudaStream_t st, int64_t n, double* A, int64_t lda, double* W, int64_t ldw, int* info) {
  if (n <= 0 || n % NB) return LFM_ERR_INVALID;
  LFM_CUDA_OK(cudaMemsetAsync(info, 0, sizeof(int), st));
  return potrf_rec(st, n, A, lda, W, ldw, info, 0);
}

// W (lower) = L^-1; diagonal 128-blocks of W already hold the leaf inverses.  The strictly upper
// block W12 of every recursion node is used as scratch for T^T = W11^T L21^T.
static int trtri_levels(cudaStream_t st, int64_t n, const double* L, int64_t ldl, double* W, int64_t ldw);

int lfm_trtri(cudaStream_t st, int64_t n, const double* L, int64_t ldl, double* W, int64_t ldw) {
  if (n == NB) return LFM_OK;
  const int64_t nblk = n / NB;
  if ((nblk & (nblk - 1)) == 0 && ldl == ldw) return trtri_levels(st, n, L, ldl, W, ldw);
  const int64_t n1 = split(n), n2 = n - n1;
  LFM_TRY(lfm_trtri(st, n1, L, ldl, W, ldw));
  LFM_TRY(lfm_trtri(st, n2, L + n1 * ldl + n1, ldl, W + n1 * ldw + n1, ldw));
  double* Tt = W + n1;              // n1 x n2 scratch (upper-right block)
  double* W21 = W + n1 * ldw;       // n2 x n1
  const double* W22 = W + n1 * ldw + n1;
  // Tt[c][i] = sum_{k >= c} W11[k][c] L21[i][k]
  LFM_TRY(lfm_dgemm(st, mk(1, 1, n1, n2, n1, W, ldw, L + n1 * ldl, ldl, Tt, ldw, 1.0, 0.0, 0, LFM_K_GE_ROW)));
  // W21[i][c] = - sum_{k <= i} W22[i][k] T[k][c],  T stored transposed (N x K)
  return lfm_dgemm(st, mk(0, 1, n2, n1, n2, W22, ldw, Tt, ldw, W21, ldw, -1.0, 0.0, 0, LFM_K_LE_ROW));
}

// Power-of-two block counts: the recursion tree is regular, so all nodes of one level (independent
// diagonal blocks, identical shapes, constant stride along the diagonal) run as ONE batched launch:
// 2 log2(n / 128) launches in total.
static int trtri_levels(cudaStream_t st, int64_t n, const double* L, int64_t ld, double* W, int64_t ldw) {
  for (int64_t m = NB; m < n; m *= 2) {  // m = half size of the nodes at this level
    const int64_t count = n / (2 * m);
    const int64_t stride = 2 * m * (ld + 1);
    LfmGemm g1 = mk(1, 1, m, m, m, W, ldw, L + m * ld, ld, W + m, ldw, 1.0, 0.0, 0, LFM_K_GE_ROW);
    g1.batch = (int)count; g1.strideA = stride; g1.strideB = stride; g1.strideC = stride;
    LFM_TRY(lfm_dgemm(st, g1));
    LfmGemm g2 = mk(0, 1, m, m, m, W + m * ldw + m, ldw, W + m, ldw, W + m * ldw, ldw, -1.0, 0.0, 0, LFM_K_LE_ROW);
    g2.batch = (int)count; g2.strideA = stride; g2.strideB = stride; g2.strideC = stride;
    LFM_TRY(lfm_dgemm(st, g2));
  }
  return LFM_OK;
}

// S (lower) = W^T W, out of place: every lower tile (i,j) is an independent TN product over the
// block rows k >= i, so the whole N^3/3 is ONE launch (W's diagonal blocks have exact zeros above
// the diagonal, and the scratch that lfm_trtri leaves in W's strict upper blocks is never read).
int lfm_lauum(cudaStream_t st, int64_t n, const double* W, int64_t ldw, double* S, int64_t lds) {
  return lfm_dgemm(st, mk(1, 0, n, n, n, W, ldw, W, ldw, S, lds, 1.0, 0.0, 1, LFM_K_GE_ROWCOL));
}
// The rest of S = W^T W when the top-left half block already holds S11' = W11^T W11 (lfm_potrf_trtri_diag, early_done):
// S11 += W21^T W21 (full k-range over the second half's rows, beta = 1) and the tiles of the block rows >= n/2 as in
// lfm_lauum.
int lfm_lauum_late(cudaStream_t st, int64_t n, const double* W, int64_t ldw, double* S, int64_t lds) {
  // ONE launch of 128 x 128 tiles, longest k-range first: the tiles of S11 (k over the second half's rows, accumulating) and
  // the tiles of the block rows >= n/2 (as in lfm_lauum) fill each other's last wave (two launches: 407 + 296 us at n = 4096)
  LfmGemm g = mk(1, 0, n, n, n, W, ldw, W, ldw, S, lds, 1.0, 0.0, 1, LFM_K_LAUUM_LATE);
  g.k_split = n / 2;
  g.tile = 2;
  return lfm_dgemm(st, g);
}

// Debug: one leaf with clock64() stamps at its phase boundaries (16 values: start, loaded, then per
// sub-block [diag, trailing] x 4, L stored, 3 inverse levels, W stored).
extern "C" int lfm_debug_leaf_profile(lfm_stream_t stream, double* A, double* W, int* info, long long* stamps) {
  LFM_CUDA_OK(lfm_ensure_smem(lfm_potrf_leaf_kernel, g_leaf_smem, LEAF_SMEM));
  lfm_potrf_leaf_kernel<<<1, LEAF_THREADS, LEAF_SMEM, (cudaStream_t)stream>>>(A, NB, W, NB, info, 0, stamps);
  LFM_CUDA_OK(cudaGetLastError());
  return LFM_OK;
}
