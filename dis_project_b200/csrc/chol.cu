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
// The 128 x 128 leaves are factorised AND inverted by one CTA in shared memory; the inverted
// diagonal blocks turn every leaf-level TRSM into a GEMM.
#include "lfm_common.cuh"

#define NB LFM_NB
#define LEAF_LD (NB + 1)
#define LEAF_SMEM (NB * LEAF_LD * 8)

// One CTA, 128 threads (thread i <-> row i / column i).
__global__ void __launch_bounds__(NB) lfm_potrf_leaf_kernel(double* __restrict__ A, int64_t lda,
                                                          double* __restrict__ W, int64_t ldw,
                                                          int* __restrict__ info, int pivot_base) {
  extern __shared__ double S[];  // [NB][LEAF_LD]; lower = L, strict upper (shifted) = W^T
  __shared__ double piv;
  const int tid = threadIdx.x;
  for (int r = 0; r < NB; ++r) S[r * LEAF_LD + tid] = (tid <= r) ? A[(int64_t)r * lda + tid] : 0.0;
  __syncthreads();
  // left-looking column Cholesky: thread i owns row i
  for (int k = 0; k < NB; ++k) {
    double v = 0.0;
    if (tid >= k) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      const double* ri = S + tid * LEAF_LD;
      const double* rk = S + k * LEAF_LD;
      int m = 0;
      for (; m + 4 <= k; m += 4) {
        s0 += ri[m] * rk[m];
        s1 += ri[m + 1] * rk[m + 1];
        s2 += ri[m + 2] * rk[m + 2];
        s3 += ri[m + 3] * rk[m + 3];
      }
      for (; m < k; ++m) s0 += ri[m] * rk[m];
      v = ri[k] - ((s0 + s1) + (s2 + s3));
      if (tid == k) piv = v;
    }
    __syncthreads();
    const double p = piv;
    if (tid == k && !(p > 0.0)) atomicCAS(info, 0, pivot_base + k + 1);
    if (tid >= k) {
      const double dk = sqrt(p);
      S[tid * LEAF_LD + k] = (tid == k) ? dk : v / dk;
    }
    __syncthreads();
  }
  // write L (zero strict upper)
  for (int r = 0; r < NB; ++r) A[(int64_t)r * lda + tid] = (tid <= r) ? S[r * LEAF_LD + tid] : 0.0;
  // inverse by forward substitution, thread c owns column c of W = L^-1.
  // W[i][c] (i > c) is kept at S[c][i + 1] (the unused strict upper part, shifted by one column).
  {
    const int c = tid;
    const double wcc = 1.0 / S[c * LEAF_LD + c];
    const int cmin = (tid >> 5) << 5;  // warp-uniform loop bounds -> broadcast reads of L
    double* wc = S + c * LEAF_LD + 1;  // wc[k] = W[k][c]
    for (int i = cmin + 1; i < NB; ++i) {
      const double* li = S + i * LEAF_LD;
      double s0 = 0.0, s1 = 0.0;
      int k = cmin;
      for (; k + 2 <= i; k += 2) {
        const double w0 = (k == c) ? wcc : wc[k];
        const double w1 = (k + 1 == c) ? wcc : wc[k + 1];
        if (k >= c) s0 += li[k] * w0;
        if (k + 1 >= c) s1 += li[k + 1] * w1;
      }
      for (; k < i; ++k) {
        const double w0 = (k == c) ? wcc : wc[k];
        if (k >= c) s0 += li[k] * w0;
      }
      if (i > c) wc[i] = -(s0 + s1) / li[i];
    }
    __syncthreads();
    for (int r = 0; r < NB; ++r) {
      double v;
      if (tid < r) v = S[tid * LEAF_LD + r + 1];
      else if (tid == r) v = 1.0 / S[r * LEAF_LD + r];
      else v = 0.0;
      W[(int64_t)r * ldw + tid] = v;
    }
  }
}

static int leaf(cudaStream_t st, double* A, int64_t lda, double* W, int64_t ldw, int* info, int64_t pivot_base) {
  static bool configured = false;
  if (!configured) {
    LFM_CUDA_OK(cudaFuncSetAttribute(lfm_potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
    configured = true;
  }
  lfm_potrf_leaf_kernel<<<1, NB, LEAF_SMEM, st>>>(A, lda, W, ldw, info, (int)pivot_base);
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

static int potrf_rec(cudaStream_t st, int64_t n, double* A, int64_t lda, double* W, int64_t ldw, int* info,
                     int64_t pivot_base) {
  if (n == NB) return leaf(st, A, lda, W, ldw, info, pivot_base);
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
int lfm_trtri(cudaStream_t st, int64_t n, const double* L, int64_t ldl, double* W, int64_t ldw) {
  if (n == NB) return LFM_OK;
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

// S (lower) = W^T W, out of place: every lower tile (i,j) is an independent TN product over the
// block rows k >= i, so the whole N^3/3 is ONE launch (W's diagonal blocks have exact zeros above
// the diagonal, and the scratch that lfm_trtri leaves in W's strict upper blocks is never read).
int lfm_lauum(cudaStream_t st, int64_t n, const double* W, int64_t ldw, double* S, int64_t lds) {
  return lfm_dgemm(st, mk(1, 0, n, n, n, W, ldw, W, ldw, S, lds, 1.0, 0.0, 1, LFM_K_GE_ROWCOL));
}
