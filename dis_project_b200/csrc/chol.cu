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
#include "lfm_common.cuh"

#define NB LFM_NB
#define LB 32                       // sub-block edge inside a leaf
#define S_LD 132                    // == 4 (mod 16): conflict-free m8n8k4 fragment reads
#define WD_LD 36                    // == 4 (mod 16)
#define LEAF_THREADS 256
#define LEAF_SMEM ((NB * S_LD + 4 * LB * WD_LD + LB * 33) * 8)

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

// Diagonal 32 x 32 sub-block: warp 0 factorises it, warp 1 inverts it one pivot behind.
// Single-warp code on an otherwise idle SM sub-partition runs at several cycles per instruction, so the
// two rank-1 update streams (row i of L on lane i of warp 0; column c of Y = L^-1 on lane c of warp 1)
// are put on different schedulers and coupled only by a progress counter in shared memory: after
// pivot k, warp 0 publishes column k of L (in the [32][33] scratch T, bank-conflict free for column
// reads) and 1/L_kk; warp 1 then computes W[k][c] = Y[k][c] / L_kk and Y[i][c] -= L[i][k] W[k][c].
#define DP_LD 33
__device__ __forceinline__ int warp_potrf32(double* __restrict__ D, int ldd, double* __restrict__ T,
                                            double* __restrict__ rdiag, volatile int* prog, int base, int lane) {
  int fail = -1;
#pragma unroll 8
  for (int r = 0; r < LB; ++r) T[r * DP_LD + lane] = D[r * ldd + lane];
  __syncwarp();
  double* myrow = T + lane * DP_LD;
  for (int k = 0; k < LB; ++k) {
    const double akk = T[k * DP_LD + k];
    if (!(akk > 0.0) && fail < 0) fail = k;
    const double rk = rsqrt(akk);
    double lik = 0.0;
    if (lane >= k) {
      lik = (lane == k) ? akk * rk : myrow[k] * rk;
      myrow[k] = lik;
      if (lane == k) rdiag[k] = rk;
    }
    __syncwarp();
    if (lane == 0) { __threadfence_block(); *prog = base + k + 1; }
    for (int j0 = k + 1; j0 < LB; j0 += 8) {
      double l[8], a[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int j = (j0 + q < LB) ? j0 + q : LB - 1;
        l[q] = T[j * DP_LD + k];  // L[j][k], broadcast
        a[q] = myrow[j];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int j = j0 + q;
        if (j < LB && lane >= j) myrow[j] = fma(-lik, l[q], a[q]);
      }
    }
    __syncwarp();
  }
#pragma unroll 8
  for (int r = 0; r < LB; ++r) D[r * ldd + lane] = (lane <= r) ? T[r * DP_LD + lane] : 0.0;
  return fail;
}

__device__ __forceinline__ void warp_trtri32(const double* __restrict__ T, double* __restrict__ Winv,
                                             const double* __restrict__ rdiag, volatile int* prog, int base,
                                             int lane) {
#pragma unroll 8
  for (int r = 0; r < LB; ++r) Winv[r * WD_LD + lane] = (r == lane) ? 1.0 : 0.0;
  for (int k = 0; k < LB; ++k) {
    while (*prog < base + k + 1) {}
    __threadfence_block();
    const double wk = Winv[k * WD_LD + lane] * rdiag[k];
    Winv[k * WD_LD + lane] = wk;
    for (int i0 = k + 1; i0 < LB; i0 += 8) {
      double l[8], y[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = (i0 + q < LB) ? i0 + q : LB - 1;
        l[q] = T[i * DP_LD + k];
        y[q] = Winv[i * WD_LD + lane];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = i0 + q;
        if (i < LB) Winv[i * WD_LD + lane] = fma(-l[q], wk, y[q]);
      }
    }
  }
}

// Leaf: in-place Cholesky of one 128 x 128 diagonal block AND its inverse, one CTA of 8 warps.
// Blocked 4 x 4 over 32 x 32 sub-blocks; all sub-block products run on DMMA from shared memory.
__global__ void __launch_bounds__(LEAF_THREADS, 1) lfm_potrf_leaf_kernel(double* __restrict__ A, int64_t lda,
                                                                       double* __restrict__ W, int64_t ldw,
                                                                       int* __restrict__ info, int pivot_base,
                                                                       long long* __restrict__ stamps) {
  extern __shared__ __align__(16) double S[];   // [NB][S_LD] then Wd[4][LB][WD_LD]
  int nstamp = 0;
#define LEAF_STAMP() do { if (stamps && threadIdx.x == 0) stamps[nstamp++] = clock64(); } while (0)
  LEAF_STAMP();
  double* Wd = S + NB * S_LD;
  __shared__ int failed;
  __shared__ int progress;
  __shared__ double rdiag[LB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { failed = -1; progress = 0; }
  // load the lower triangle (coalesced 128-double rows), zero above the diagonal
#pragma unroll 8
  for (int idx = tid; idx < NB * NB; idx += LEAF_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    S[r * S_LD + c] = (c <= r) ? A[(int64_t)r * lda + c] : 0.0;
  }
  __syncthreads();
  LEAF_STAMP();
  for (int kb = 0; kb < 4; ++kb) {
    double* Dkk = S + (kb * LB) * S_LD + kb * LB;
    if (warp == 0) {
      const int f = warp_potrf32(Dkk, S_LD, Wd + 4 * LB * WD_LD, rdiag, &progress, kb * LB, lane);
      if (lane == 0 && f >= 0 && failed < 0) failed = kb * LB + f;
    } else if (warp == 1) {
      warp_trtri32(Wd + 4 * LB * WD_LD, Wd + kb * LB * WD_LD, rdiag, &progress, kb * LB, lane);
    }
    __syncthreads();
    LEAF_STAMP();
    // panel: S[ib][kb] <- S[ib][kb] * Winv_kk^T
    if (warp >= 1 && kb + warp < 4) {
      double* C = S + ((kb + warp) * LB) * S_LD + kb * LB;
      double acc[4][4][2];
      warp_zero32(acc);
      warp_gemm32<1>(acc, C, S_LD, Wd + kb * LB * WD_LD, WD_LD, lane);
      __syncwarp();
      warp_store32(acc, C, S_LD, 1.0, 0.0, lane);
    }
    __syncthreads();
    // trailing update: S[ib][jb] -= S[ib][kb] S[jb][kb]^T for ib >= jb > kb
    {
      int cnt = 0;
      for (int ib = kb + 1; ib < 4; ++ib)
        for (int jb = kb + 1; jb <= ib; ++jb, ++cnt)
          if (cnt == warp) {
            double acc[4][4][2];
            warp_zero32(acc);
            warp_gemm32<1>(acc, S + (ib * LB) * S_LD + kb * LB, S_LD, S + (jb * LB) * S_LD + kb * LB, S_LD, lane);
            warp_store32(acc, S + (ib * LB) * S_LD + jb * LB, S_LD, -1.0, 1.0, lane);
          }
    }
    __syncthreads();
    LEAF_STAMP();
  }
  // L out (strict upper blocks are still the zeros of the load)
#pragma unroll 8
  for (int idx = tid; idx < NB * NB; idx += LEAF_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    A[(int64_t)r * lda + c] = S[r * S_LD + c];
  }
  LEAF_STAMP();
  // W = L^-1 by block forward substitution; W_ij (i > j) is built in the free upper block (j, i):
  //   W_ij = -Winv_ii * sum_{k=j}^{i-1} L_ik W_kj
  for (int d = 1; d < 4; ++d) {
    const int j = warp, i = warp + d;
    if (i < 4) {
      double* dst = S + (j * LB) * S_LD + i * LB;
      double acc[4][4][2];
      warp_zero32(acc);
      for (int k = j; k < i; ++k) {
        const double* Lik = S + (i * LB) * S_LD + k * LB;
        if (k == j) warp_gemm32<0>(acc, Lik, S_LD, Wd + j * LB * WD_LD, WD_LD, lane);
        else warp_gemm32<0>(acc, Lik, S_LD, S + (j * LB) * S_LD + k * LB, S_LD, lane);
      }
      warp_store32(acc, dst, S_LD, 1.0, 0.0, lane);
      __syncwarp();
      warp_zero32(acc);
      warp_gemm32<0>(acc, Wd + i * LB * WD_LD, WD_LD, dst, S_LD, lane);
      __syncwarp();
      warp_store32(acc, dst, S_LD, -1.0, 0.0, lane);
    }
    __syncthreads();
    LEAF_STAMP();
  }
#pragma unroll 8
  for (int idx = tid; idx < NB * NB; idx += LEAF_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    const int bi = r >> 5, bj = c >> 5;
    double v = 0.0;
    if (bi == bj) v = Wd[bi * LB * WD_LD + (r & 31) * WD_LD + (c & 31)];
    else if (bi > bj) v = S[(bj * LB + (r & 31)) * S_LD + bi * LB + (c & 31)];
    W[(int64_t)r * ldw + c] = v;
  }
  LEAF_STAMP();
  if (tid == 0 && failed >= 0) atomicCAS(info, 0, pivot_base + failed + 1);
}

static int leaf(cudaStream_t st, double* A, int64_t lda, double* W, int64_t ldw, int* info, int64_t pivot_base) {
  static bool configured = false;
  if (!configured) {
    LFM_CUDA_OK(cudaFuncSetAttribute(lfm_potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
    configured = true;
  }
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

// Right-looking blocked Cholesky with 128-wide panels for small / medium n: every step is a leaf, one
// in-place panel multiply by the inverted diagonal block and one fat trailing SYRK (3 launches per
// block column instead of the O(log) tiny TRSM launches of the recursion).
static int potrf_right_looking(cudaStream_t st, int64_t n, double* A, int64_t lda, double* W, int64_t ldw, int* info,
                               int64_t pivot_base) {
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

// Debug: one leaf with clock64() stamps at its phase boundaries (16 values: start, loaded, then per
// sub-block [diag, trailing] x 4, L stored, 3 inverse levels, W stored).
extern "C" int lfm_debug_leaf_profile(lfm_stream_t stream, double* A, double* W, int* info, long long* stamps) {
  static bool configured = false;
  if (!configured) {
    LFM_CUDA_OK(cudaFuncSetAttribute(lfm_potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
    configured = true;
  }
  lfm_potrf_leaf_kernel<<<1, LEAF_THREADS, LEAF_SMEM, (cudaStream_t)stream>>>(A, NB, W, NB, info, 0, stamps);
  LFM_CUDA_OK(cudaGetLastError());
  return LFM_OK;
}
