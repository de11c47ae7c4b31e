// Latent posterior mean / variance at T* test inputs (ExactLFM.latent_predict, src/model.py:420-463).
//
//   Sigma_p = K + diag(variances) + jitter I = L L^T,  W = L^-1,  alpha = Sigma_p^-1 (y - mean_x)
//   mean_i  = mean_t,i + k_i^T alpha                      (model.py:452-454)
//   var_i   = k(t*_i, t*_i) + 2 jitter - || W k_i ||^2     (model.py:456-461; jitter twice)
//
// The N x T* cross-covariance is never held whole: test inputs are streamed in chunks of PC_COLS
// columns -- build the K_xf chunk, one triangular DMMA GEMM V = W K_xf, then column reductions.
#include "sim_math.cuh"


#define PC_COLS 2048
#define PR_ROWS 1024

int lfm_launch_residual(cudaStream_t st, int64_t N, int64_t Npad, const double* X, const double* y, int G,
                        const double* theta, double* z, double* out_mean);
size_t lfm_alpha_part_doubles(int64_t Np);
int lfm_launch_alpha(cudaStream_t st, int64_t Np, const double* W, const double* z, double* w, double* part,
                     double* alpha);

// zero rows [N, Npad) of an Npad x cols row-major block
__global__ void lfm_zero_rows_kernel(double* __restrict__ A, int64_t ld, int64_t r0, int64_t r1, int64_t cols) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nr = r1 - r0;
  if (idx >= nr * cols) return;
  A[(r0 + idx / cols) * ld + idx % cols] = 0.0;
}

// partial column reductions over a row chunk: pm[chunk][j] = sum_i Kxf[i][j] alpha[i],
// pq[chunk][j] = sum_i V[i][j]^2.  128 columns per CTA.
__global__ void __launch_bounds__(128) lfm_post_colred_kernel(int64_t n, int64_t cols, const double* __restrict__ Kxf,
                                                            const double* __restrict__ V, int64_t ld,
                                                            const double* __restrict__ alpha,
                                                            double* __restrict__ pm, double* __restrict__ pq) {
  const int64_t j = blockIdx.x * 128 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * PR_ROWS;
  const int64_t r1 = min(n, r0 + PR_ROWS);
  double m0 = 0.0, m1 = 0.0, q0 = 0.0, q1 = 0.0;
  int64_t i = r0;
  for (; i + 2 <= r1; i += 2) {
    const double k0 = Kxf[i * ld + j], k1 = Kxf[(i + 1) * ld + j];
    const double v0 = V[i * ld + j], v1 = V[(i + 1) * ld + j];
    m0 += k0 * alpha[i]; m1 += k1 * alpha[i + 1];
    q0 += v0 * v0; q1 += v1 * v1;
  }
  for (; i < r1; ++i) {
    m0 += Kxf[i * ld + j] * alpha[i];
    const double v0 = V[i * ld + j];
    q0 += v0 * v0;
  }
  pm[(int64_t)blockIdx.y * cols + j] = m0 + m1;
  pq[(int64_t)blockIdx.y * cols + j] = q0 + q1;
}

// final: mean/var for the chunk's test points
__global__ void lfm_post_finish_kernel(int64_t ncols, int64_t cols_ld, int nchunk, const double* __restrict__ pm,
                                       const double* __restrict__ pq, const double* __restrict__ Xstar,
                                       int64_t t0, int64_t Tstar, int G, const double* __restrict__ theta,
                                       double jitter, const int* __restrict__ info, double* __restrict__ out_mean,
                                       double* __restrict__ out_var) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  double m = 0.0, q = 0.0;
  for (int k = 0; k < nchunk; ++k) { m += pm[(int64_t)k * cols_ld + c]; q += pq[(int64_t)k * cols_ld + c]; }
  const int64_t i = t0 + c;
  const double l = theta[3 * G];
  const LfmPoint p = lfm_make_point(Xstar + 3 * i, G, theta, theta + G, l, false);
  const double kdiag = lfm_kernel(p, p, l, 1.0 / l);
  int64_t block = Tstar / G;
  if (block < 1) block = 1;
  int64_t g = i / block;
  if (g > G - 1) g = G - 1;
  const double mean_t = theta[2 * G + g] / theta[g] * (double)p.flag;  // model.py:444
  double mean = mean_t + m;
  double var = kdiag + jitter - q + jitter;
  if (*info != 0) { mean = nan(""); var = nan(""); }
  out_mean[i] = mean;
  out_var[i] = var;
}

struct PostWs {
  int64_t Np;
  double *A, *W, *z, *w, *alpha, *part, *Kxf, *V, *pm, *pq;
  size_t total_doubles;
};
static PostWs post_ws_layout(int64_t N, void* base) {
  PostWs s;
  s.Np = lfm_round_up(N, LFM_NB);
  double* p = reinterpret_cast<double*>(base);
  size_t off = 0;
  auto take = [&](size_t n) { double* r = p ? p + off : nullptr; off += (n + 1) & ~(size_t)1; return r; };
  s.A = take((size_t)s.Np * s.Np);
  s.W = take((size_t)s.Np * s.Np);
  s.z = take(s.Np);
  s.w = take(s.Np);
  s.alpha = take(s.Np);
  s.part = take(lfm_alpha_part_doubles(s.Np));
  s.Kxf = take((size_t)s.Np * PC_COLS);
  s.V = take((size_t)s.Np * PC_COLS);
  const size_t nchunk = (size_t)((s.Np + PR_ROWS - 1) / PR_ROWS);
  s.pm = take(nchunk * PC_COLS);
  s.pq = take(nchunk * PC_COLS);
  s.total_doubles = off;
  return s;
}

extern "C" size_t lfm_latent_posterior_workspace_bytes(int64_t N, int G, int64_t Tstar) {
  (void)G; (void)Tstar;
  if (N <= 0) return 0;
  return post_ws_layout(N, nullptr).total_doubles * sizeof(double);
}

extern "C" int lfm_latent_posterior(lfm_stream_t stream, int64_t N, int G, const double* X, const double* y,
                                    const double* variances, const double* theta, double jitter, int64_t Tstar,
                                    const double* Xstar, void* ws, size_t ws_bytes, double* out_mean,
                                    double* out_var, int* info) {
  if (N <= 0 || G <= 0 || Tstar <= 0 || !X || !y || !variances || !theta || !Xstar || !ws || !out_mean || !out_var ||
      !info)
    return LFM_ERR_INVALID;
  if (N % G) return LFM_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return LFM_ERR_INVALID;
  if (ws_bytes < lfm_latent_posterior_workspace_bytes(N, G, Tstar)) return LFM_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const PostWs s = post_ws_layout(N, ws);
  const int64_t Np = s.Np;
  // factorisation of Sigma_p and alpha
  LFM_TRY(lfm_launch_residual(st, N, Np, X, y, G, theta, s.z, nullptr));
  LFM_TRY(lfm_launch_sigma_lower(st, N, Np, X, G, theta, variances, jitter, 0, s.A, Np));
  LFM_TRY(lfm_potrf_trtri(st, Np, s.A, Np, s.W, Np, info));
  LFM_TRY(lfm_launch_alpha(st, Np, s.W, s.z, s.w, s.part, s.alpha));
  // stream the test inputs
  const int nrch = (int)((Np + PR_ROWS - 1) / PR_ROWS);
  for (int64_t t0 = 0; t0 < Tstar; t0 += PC_COLS) {
    const int64_t nc = (Tstar - t0 < PC_COLS) ? (Tstar - t0) : PC_COLS;
    const int64_t ncp = lfm_round_up(nc, 128);
    if (ncp != nc || Np != N) {
      // pad columns / rows of the chunk must be finite zeros for the GEMM
      LFM_CUDA_OK(cudaMemsetAsync(s.Kxf, 0, sizeof(double) * (size_t)Np * PC_COLS, st));
    }
    LFM_TRY(lfm_launch_cross_cov(st, N, nc, X, Xstar + 3 * t0, G, theta, s.Kxf, PC_COLS));
    LfmGemm g;
    g.transA = 0; g.transB = 0; g.M = Np; g.N = ncp; g.K = Np;
    g.A = s.W; g.lda = Np; g.B = s.Kxf; g.ldb = PC_COLS; g.C = s.V; g.ldc = PC_COLS;
    g.alpha = 1.0; g.beta = 0.0; g.lower_only = 0; g.kmode = LFM_K_LE_ROW;
    g.batch = 1; g.strideA = g.strideB = g.strideC = 0;
    LFM_TRY(lfm_dgemm(st, g));
    lfm_post_colred_kernel<<<dim3((unsigned)(ncp / 128), (unsigned)nrch), 128, 0, st>>>(Np, PC_COLS, s.Kxf, s.V,
                                                                                      PC_COLS, s.alpha, s.pm, s.pq);
    LFM_LAUNCHED(1);
    LFM_CUDA_OK(cudaGetLastError());
    lfm_post_finish_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(nc, PC_COLS, nrch, s.pm, s.pq, Xstar, t0,
                                                                        Tstar, G, theta, jitter, info, out_mean,
                                                                        out_var);
    LFM_LAUNCHED(1);
    LFM_CUDA_OK(cudaGetLastError());
  }
  return LFM_OK;
}

// =================================================================================================
// Gene-expression posterior (ExactLFM.multi_gene_predict, src/model.py:465-514): third noise model
//   Sigma_g = K + diag(variances) + sigma^2 I            (no jitter)
//   mean = mean_t + K_tx Sigma_g^-1 (y - mean_x),  cov = K_tt - K_tx Sigma_g^-1 K_xt + jitter I
// Test rows go through the general flag-aware kernel (gene indices follow jnp indexing).
// cov = K_tt - V^T V with V = W K_xt: one triangular DMMA GEMM + one TN SYRK on the lower tiles.
// =================================================================================================
__global__ void lfm_gene_finish_kernel(int64_t T, int64_t Tp, const double* __restrict__ Cg,
                                       double* __restrict__ cov, double* __restrict__ var,
                                       const double* __restrict__ Xstar, int G, const double* __restrict__ theta,
                                       double jitter, const int* __restrict__ info) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= T) return;
  const bool bad = *info != 0;
  const double g = (i >= j) ? Cg[i * Tp + j] : Cg[j * Tp + i];
  if (cov) {
    double v = cov[i * T + j] - g + ((i == j) ? jitter : 0.0);  // cov holds K_tt on entry
    if (bad) v = nan("");
    cov[i * T + j] = v;
    if (i == j && var) var[i] = v;
  } else if (i == j && var) {
    const double l = theta[3 * G];
    const LfmPoint p = lfm_make_point(Xstar + 3 * i, G, theta, theta + G, l, false);
    double v = lfm_kernel(p, p, l, 1.0 / l) - g + jitter;
    if (bad) v = nan("");
    var[i] = v;
  }
}
__global__ void lfm_gene_mean_kernel(int64_t T, int64_t cols_ld, int nchunk, const double* __restrict__ pm,
                                     const double* __restrict__ Xstar, int G, const double* __restrict__ theta,
                                     const int* __restrict__ info, double* __restrict__ out_mean) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= T) return;
  double m = 0.0;
  for (int k = 0; k < nchunk; ++k) m += pm[(int64_t)k * cols_ld + c];
  int64_t block = T / G;
  if (block < 1) block = 1;
  int64_t g = c / block;
  if (g > G - 1) g = G - 1;
  const double flag = (double)((int)Xstar[3 * c + 2]);
  double mean = theta[2 * G + g] / theta[g] * flag + m;  // model.py:501,507
  if (*info != 0) mean = nan("");
  out_mean[c] = mean;
}

struct GeneWs {
  int64_t Np, Tp;
  double *A, *W, *z, *w, *alpha, *part, *Kxt, *V, *Cg, *pm, *pq;
  size_t total_doubles;
};
static GeneWs gene_ws_layout(int64_t N, int64_t T, void* base) {
  GeneWs s;
  s.Np = lfm_round_up(N, LFM_NB);
  s.Tp = lfm_round_up(T, LFM_NB);
  double* p = reinterpret_cast<double*>(base);
  size_t off = 0;
  auto take = [&](size_t n) { double* r = p ? p + off : nullptr; off += (n + 1) & ~(size_t)1; return r; };
  s.A = take((size_t)s.Np * s.Np);
  s.W = take((size_t)s.Np * s.Np);
  s.z = take(s.Np);
  s.w = take(s.Np);
  s.alpha = take(s.Np);
  s.part = take(lfm_alpha_part_doubles(s.Np));
  s.Kxt = take((size_t)s.Np * s.Tp);
  s.V = take((size_t)s.Np * s.Tp);
  s.Cg = take((size_t)s.Tp * s.Tp);
  const size_t nchunk = (size_t)((s.Np + PR_ROWS - 1) / PR_ROWS);
  s.pm = take(nchunk * s.Tp);
  s.pq = take(nchunk * s.Tp);
  s.total_doubles = off;
  return s;
}

extern "C" size_t lfm_gene_posterior_workspace_bytes(int64_t N, int G, int64_t Tstar) {
  (void)G;
  if (N <= 0 || Tstar <= 0) return 0;
  return gene_ws_layout(N, Tstar, nullptr).total_doubles * sizeof(double);
}

extern "C" int lfm_gene_posterior(lfm_stream_t stream, int64_t N, int G, const double* X, const double* y,
                                  const double* variances, const double* theta, double jitter, int64_t Tstar,
                                  const double* Xstar, void* ws, size_t ws_bytes, double* out_mean, double* out_cov,
                                  double* out_var, int* info) {
  if (N <= 0 || G <= 0 || Tstar <= 0 || !X || !y || !variances || !theta || !Xstar || !ws || !out_mean || !info)
    return LFM_ERR_INVALID;
  if (N % G || Tstar % G) return LFM_ERR_INVALID;  // mean_function reshape (model.py:145-149)
  if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return LFM_ERR_INVALID;
  if (ws_bytes < lfm_gene_posterior_workspace_bytes(N, G, Tstar)) return LFM_ERR_WORKSPACE;
  if (Tstar > 65535) return LFM_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const GeneWs s = gene_ws_layout(N, Tstar, ws);
  const int64_t Np = s.Np, Tp = s.Tp;
  LFM_TRY(lfm_launch_residual(st, N, Np, X, y, G, theta, s.z, nullptr));
  LFM_TRY(lfm_launch_sigma_lower(st, N, Np, X, G, theta, variances, 0.0, 1, s.A, Np));
  LFM_TRY(lfm_potrf_trtri(st, Np, s.A, Np, s.W, Np, info));
  LFM_TRY(lfm_launch_alpha(st, Np, s.W, s.z, s.w, s.part, s.alpha));
  LFM_CUDA_OK(cudaMemsetAsync(s.Kxt, 0, sizeof(double) * (size_t)Np * Tp, st));
  LFM_TRY(lfm_launch_cross_cov(st, N, Tstar, X, Xstar, G, theta, s.Kxt, Tp));
  LfmGemm g;
  g.transA = 0; g.transB = 0; g.M = Np; g.N = Tp; g.K = Np;
  g.A = s.W; g.lda = Np; g.B = s.Kxt; g.ldb = Tp; g.C = s.V; g.ldc = Tp;
  g.alpha = 1.0; g.beta = 0.0; g.lower_only = 0; g.kmode = LFM_K_LE_ROW;
  g.batch = 1; g.strideA = g.strideB = g.strideC = 0;
  LFM_TRY(lfm_dgemm(st, g));
  const int nrch = (int)((Np + PR_ROWS - 1) / PR_ROWS);
  lfm_post_colred_kernel<<<dim3((unsigned)(Tp / 128), (unsigned)nrch), 128, 0, st>>>(Np, Tp, s.Kxt, s.V, Tp, s.alpha,
                                                                                    s.pm, s.pq);
  LFM_LAUNCHED(1);
  LFM_CUDA_OK(cudaGetLastError());
  lfm_gene_mean_kernel<<<(unsigned)((Tstar + 255) / 256), 256, 0, st>>>(Tstar, Tp, nrch, s.pm, Xstar, G, theta, info,
                                                                       out_mean);
  LFM_LAUNCHED(1);
  LFM_CUDA_OK(cudaGetLastError());
  // Cg (lower tiles) = V^T V
  LfmGemm c;
  c.transA = 1; c.transB = 0; c.M = Tp; c.N = Tp; c.K = Np;
  c.A = s.V; c.lda = Tp; c.B = s.V; c.ldb = Tp; c.C = s.Cg; c.ldc = Tp;
  c.alpha = 1.0; c.beta = 0.0; c.lower_only = 1; c.kmode = LFM_K_FULL;
  c.batch = 1; c.strideA = c.strideB = c.strideC = 0;
  LFM_TRY(lfm_dgemm(st, c));
  if (out_cov) LFM_TRY(lfm_launch_cross_cov(st, Tstar, Tstar, Xstar, Xstar, G, theta, out_cov, Tstar));
  if (out_cov || out_var) {
    lfm_gene_finish_kernel<<<dim3((unsigned)((Tstar + 255) / 256), (unsigned)Tstar), 256, 0, st>>>(
        Tstar, Tp, s.Cg, out_cov, out_var, Xstar, G, theta, jitter, info);
    LFM_LAUNCHED(1);
    LFM_CUDA_OK(cudaGetLastError());
  }
  return LFM_OK;
}
