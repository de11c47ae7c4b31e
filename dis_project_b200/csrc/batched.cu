// Batched small-N path: many independent LFMs (multi-start restarts, genes x replicas x candidate TFs)
// per GPU, one CTA per LFM, the whole problem resident in shared memory for all optimiser steps.
//
// Per step (JaxTrainer.step, src/trainer.py:105-131, inside the scan of :201-216):
//   theta = constrain(u)                                   (trainer.py:103)
//   Sigma = k_xx(X,X) + (jitter + sigma^2) I = L L^T        (objectives.py:70-73)
//   W = L^-1 (columns in parallel), Sigma^-1 = W^T W (entries in parallel), alpha = W^T W z
//   NLML = 1/2 [N log 2pi + 2 sum log L_ii + |W z|^2]       (objectives.py:76-78)
//   grad = sum_ab K_bar_ab dK_ab/dtheta, K_bar = 1/2 (Sigma^-1 - alpha alpha^T)   (AD of trainer.py:126)
//   u <- adam(u, grad * dtheta/du); "fix p21" hook          (trainer.py:127-128, 133-160, 205-210)
//
// Shared-memory matrix S[N][ld] is used in three roles over a step: lower = Sigma -> L -> Sigma^-1 ->
// K_bar-weighted row-gene derivative; upper = W^T -> column-gene derivative.  The sensitivity gradient
// uses the identity diag(K_bar K) = 1/2 (1 - c Sinv_aa - alpha_a z_a + c alpha_a^2), c = jitter + sigma^2,
// so no second copy of K is needed.  Every reduction has a fixed order: results are bit-reproducible.
#include "sim_math.cuh"

#define BT 256  // threads per CTA

struct BatchedArgs {
  int64_t B;
  int N, G;
  const double* X;
  const double* y;
  double* u_io;       // B x P unconstrained (in/out)
  double* adam;       // B x 2P (m, v) or NULL
  double jitter, lr, b1, b2, eps;
  int first_step, steps, total_steps, fix_params, steps_per_epoch;
  double* hist; int64_t ld_hist;
  double* theta_out;  // B x P constrained result written when the last step of the fit is reached (or NULL)
  double* eval_val;   // eval-only mode: B
  double* eval_grad;  // eval-only mode: B x P
  int* info;
};

__device__ __forceinline__ double block_sum(double v, double* red) {
  // fixed-order block reduction; all BT threads must call
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < BT / 32; ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(BT) lfm_batched_kernel(BatchedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.N, G = a.G, P = 3 * G + 2;
  const int ld = N | 1;
  const int tid = threadIdx.x;
  const int64_t bidx = blockIdx.x;
  // carve shared memory
  double* S = reinterpret_cast<double*>(smem_raw);
  double* z = S + (size_t)N * ld;
  double* w = z + N;
  double* alpha = w + N;
  double* wdiag = alpha + N;   // 1 / L_ii
  double* sdiag = wdiag + N;   // Sinv_ii
  double* dsum = sdiag + N;    // diagonal derivative term / per-point totals
  double* th = dsum + N;       // P constrained
  double* u = th + P;          // P unconstrained
  double* gr = u + P;          // P gradient (constrained, then unconstrained)
  double* am = gr + P;         // P adam m
  double* av = am + P;         // P adam v
  double* red = av + P;        // 8 + 2
  LfmPoint* pts = reinterpret_cast<LfmPoint*>(red + 16);
  __shared__ double piv;
  __shared__ int fail;

  if (tid < P) {
    u[tid] = a.u_io[bidx * P + tid];
    const bool have = a.adam != nullptr && a.first_step > 0;
    am[tid] = have ? a.adam[bidx * 2 * P + tid] : 0.0;
    av[tid] = have ? a.adam[bidx * 2 * P + P + tid] : 0.0;
  }
  if (tid == 0) fail = 0;
  __syncthreads();

  const int npairs = N * (N + 1) / 2;
  const bool eval_only = a.eval_val != nullptr;
  const int nsteps = eval_only ? 1 : a.steps;

  for (int sidx = 0; sidx < nsteps; ++sidx) {
    const int step = a.first_step + sidx;
    // ---- A. constrain -----------------------------------------------------------------------
    if (tid < P) th[tid] = (tid == 3 * G) ? lfm_l_forward(u[tid]) : lfm_softplus(u[tid]);
    __syncthreads();
    const double l = th[3 * G], inv_l = 1.0 / l, sigma = th[3 * G + 1];
    const double cdiag = a.jitter + sigma * sigma;
    // ---- B. points, residual -----------------------------------------------------------------
    if (tid < N) {
      pts[tid] = lfm_make_point(a.X + 3 * tid, G, th, th + G, l, true);
      int block = N / G;
      int m = tid / block;
      if (m > G - 1) m = G - 1;
      z[tid] = a.y[tid] - th[2 * G + m] / th[m] * (double)((int)a.X[3 * tid + 2]);
    }
    __syncthreads();
    // ---- C. Sigma (lower + diagonal) ------------------------------------------------------------
    for (int p = tid; p < npairs; p += BT) {
      int r = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
      while ((r + 1) * (r + 2) / 2 <= p) ++r;
      while (r * (r + 1) / 2 > p) --r;
      const int c = p - r * (r + 1) / 2;
      double k = lfm_kxx(pts[r], pts[c], l, inv_l);
      if (r == c) k += cdiag;
      S[r * ld + c] = k;
    }
    __syncthreads();
    // ---- D. Cholesky, left-looking, two threads per row -----------------------------------------
    {
      const int row = tid >> 1, half = tid & 1;
      for (int k = 0; k < N; ++k) {
        double v = 0.0;
        if (row >= k && row < N) {
          const double* ri = S + row * ld;
          const double* rk = S + k * ld;
          double s0 = 0.0, s1 = 0.0;
          int m = half;
          for (; m + 2 < k; m += 4) { s0 += ri[m] * rk[m]; s1 += ri[m + 2] * rk[m + 2]; }
          for (; m < k; m += 2) s0 += ri[m] * rk[m];
          v = s0 + s1;
        }
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (row >= k && row < N) {
          v = S[row * ld + k] - v;
          if (row == k && half == 0) piv = v;
        }
        __syncthreads();
        const double p = piv;
        if (tid == 0 && !(p > 0.0) && fail == 0) fail = k + 1;
        if (row >= k && row < N && half == 0) {
          const double dk = sqrt(p);
          S[row * ld + k] = (row == k) ? dk : v / dk;
        }
        __syncthreads();
      }
    }
    // ---- E. log det, W = L^-1 into the upper triangle (thread c owns column c) -------------------
    double logdet_part = 0.0;
    if (tid < N) {
      const double lii = S[tid * ld + tid];
      logdet_part = log(lii);
      wdiag[tid] = 1.0 / lii;
    }
    const double logdet = 2.0 * block_sum(logdet_part, red);
    if (tid < N) {
      const int c = tid;
      const double wcc = wdiag[c];
      const int cmin = (tid >> 5) << 5;
      double* wc = S + c * ld;  // wc[i] = W[i][c] for i > c (row c of the upper triangle)
      for (int i = cmin + 1; i < N; ++i) {
        const double* li = S + i * ld;
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
        if (i > c) wc[i] = -(s0 + s1) * wdiag[i];
      }
    }
    __syncthreads();
    // ---- F. w = W z, alpha = W^T w ------------------------------------------------------------------
    double quad_part = 0.0;
    if (tid < N) {
      double acc = wdiag[tid] * z[tid];
      for (int k = 0; k < tid; ++k) acc += S[k * ld + tid] * z[k];
      w[tid] = acc;
      quad_part = acc * acc;
    }
    const double quad = block_sum(quad_part, red);
    if (tid < N) {
      double acc = wdiag[tid] * w[tid];
      const double* uj = S + tid * ld;
      for (int i = tid + 1; i < N; ++i) acc += uj[i] * w[i];
      alpha[tid] = acc;
    }
    const double nlml = 0.5 * ((double)N * LFM_LOG_2PI + logdet + quad);
    // ---- G. Sigma^-1 = W^T W into the lower triangle (+ sdiag), every entry independent ------------
    for (int p = tid; p < npairs; p += BT) {
      int r = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
      while ((r + 1) * (r + 2) / 2 <= p) ++r;
      while (r * (r + 1) / 2 > p) --r;
      const int c = p - r * (r + 1) / 2;
      // Sinv[r][c] = sum_{k >= r} W[k][r] W[k][c];  W[k][r] = S[r][k] (k > r), W[r][r] = wdiag[r]
      const double* ur = S + r * ld;
      const double* uc = S + c * ld;
      double s0 = wdiag[r] * ((r == c) ? wdiag[r] : uc[r]);
      double s1 = 0.0;
      int k = r + 1;
      for (; k + 2 <= N; k += 2) { s0 += ur[k] * uc[k]; s1 += ur[k + 1] * uc[k + 1]; }
      for (; k < N; ++k) s0 += ur[k] * uc[k];
      const double v = s0 + s1;
      if (r == c) sdiag[r] = v;
      else S[r * ld + c] = v;  // lower; L is dead
    }
    __syncthreads();
    // ---- H/I. fused derivative contraction over the lower triangle ----------------------------------
    double dl_part = 0.0;
    for (int p = tid; p < npairs; p += BT) {
      int r = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
      while ((r + 1) * (r + 2) / 2 <= p) ++r;
      while (r * (r + 1) / 2 > p) --r;
      const int c = p - r * (r + 1) / 2;
      const double sinv = (r == c) ? sdiag[r] : S[r * ld + c];
      const double wgt = ((r == c) ? 0.5 : 1.0) * (sinv - alpha[r] * alpha[c]);
      double k, dr, dc, dl;
      lfm_kxx_grad(pts[r], pts[c], l, inv_l, k, dr, dc, dl);
      dl_part += wgt * dl;
      if (r == c) dsum[r] = wgt * (dr + dc);
      else { S[r * ld + c] = wgt * dr; S[c * ld + r] = wgt * dc; }
    }
    const double gl = block_sum(dl_part, red);
    // per-point totals: full row sums
    if (tid < N) {
      const double* rp = S + tid * ld;
      double s0 = dsum[tid], s1 = 0.0;
      int c = 0;
      for (; c + 2 <= N; c += 2) {
        s0 += (c == tid) ? 0.0 : rp[c];
        s1 += (c + 1 == tid) ? 0.0 : rp[c + 1];
      }
      for (; c < N; ++c) s0 += (c == tid) ? 0.0 : rp[c];
      dsum[tid] = s0 + s1;
    }
    __syncthreads();
    // ---- J. fold by gene; mean-function terms; sigma ---------------------------------------------
    if (tid < G) {
      const int m = tid;
      double gd = 0.0, gs = 0.0, asum = 0.0;
      for (int i = 0; i < N; ++i) {
        if (pts[i].gene == m) {
          gd += dsum[i];
          gs += 1.0 - cdiag * sdiag[i] - alpha[i] * z[i] + cdiag * alpha[i] * alpha[i];
        }
      }
      const int block = N / G;
      for (int i = m * block; i < (m + 1) * block; ++i) asum += alpha[i];
      const double D = th[m], Sm = th[G + m], Bm = th[2 * G + m];
      gr[m] = gd + asum * Bm / (D * D);
      gr[G + m] = gs / Sm;
      gr[2 * G + m] = -asum / D;
    }
    if (tid == G) {
      double tr = 0.0, aa = 0.0;
      for (int i = 0; i < N; ++i) { tr += sdiag[i]; aa += alpha[i] * alpha[i]; }
      gr[3 * G] = gl;
      gr[3 * G + 1] = sigma * (tr - aa);
    }
    __syncthreads();
    // ---- K. chain rule, Adam, hook ----------------------------------------------------------------
    const bool bad = fail != 0;
    if (tid < P) {
      const double sg = lfm_sigmoid(u[tid]);
      const double jac = (tid == 3 * G) ? (LFM_L_HIGH - LFM_L_LOW) * sg * (1.0 - sg) : sg;
      double g = gr[tid] * jac;
      if (bad) g = nan("");
      if (eval_only) {
        a.eval_grad[bidx * P + tid] = g;
      } else {
        const double m1 = a.b1 * am[tid] + (1.0 - a.b1) * g;
        const double v1 = a.b2 * av[tid] + (1.0 - a.b2) * g * g;
        am[tid] = m1; av[tid] = v1;
        const double mhat = m1 / (1.0 - pow(a.b1, (double)(step + 1)));
        const double vhat = v1 / (1.0 - pow(a.b2, (double)(step + 1)));
        double un = u[tid] - a.lr * mhat / (sqrt(vhat) + a.eps);
        if (a.fix_params && (step % a.steps_per_epoch) == 0 && G > 3) {
          if (tid == G + 3) un = 1.0;  // true_s[3]  (trainer.py:152, unconstrained space: SURVEY Q5)
          if (tid == 3) un = 0.8;      // true_d[3]  (trainer.py:153)
        }
        u[tid] = un;
      }
    }
    if (tid == 0) {
      const double v = bad ? nan("") : nlml;
      if (eval_only) a.eval_val[bidx] = v;
      else if (a.hist) a.hist[bidx * a.ld_hist + step] = v;
    }
    __syncthreads();
  }

  if (!eval_only && tid < P) {
    a.u_io[bidx * P + tid] = u[tid];
    if (a.adam) { a.adam[bidx * 2 * P + tid] = am[tid]; a.adam[bidx * 2 * P + P + tid] = av[tid]; }
    if (a.theta_out && a.first_step + a.steps >= a.total_steps) {
      double t = (tid == 3 * G) ? lfm_l_forward(u[tid]) : lfm_softplus(u[tid]);  // trainer.py:218
      if (a.fix_params && G > 3) {                                               // trainer.py:219-220
        if (tid == G + 3) t = 1.0;
        if (tid == 3) t = 0.8;
      }
      a.theta_out[bidx * P + tid] = t;
    }
  }
  if (tid == 0 && a.info) {
    if (a.first_step == 0 || eval_only) a.info[bidx] = fail;
    else if (fail) a.info[bidx] = fail;
  }
}

static size_t batched_smem_bytes(int N, int G) {
  const size_t P = 3 * (size_t)G + 2;
  const size_t ld = (size_t)(N | 1);
  size_t d = (size_t)N * ld + 6 * (size_t)N + 5 * P + 16;
  return d * 8 + (size_t)N * sizeof(LfmPoint);
}

static int batched_launch(cudaStream_t st, const BatchedArgs& a) {
  if (a.B <= 0 || a.N <= 0 || a.G <= 0 || !a.X || !a.y || !a.u_io) return LFM_ERR_INVALID;
  if (a.N % a.G) return LFM_ERR_INVALID;
  if (a.N > 128 || 3 * a.G + 2 > BT || a.B > 0x7fffffff) return LFM_ERR_UNSUPPORTED;
  const size_t smem = batched_smem_bytes(a.N, a.G);
  if (smem > 227 * 1024) return LFM_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    LFM_CUDA_OK(cudaFuncSetAttribute(lfm_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  lfm_batched_kernel<<<(unsigned)a.B, BT, smem, st>>>(a);
  LFM_LAUNCHED(1);
  LFM_CUDA_OK(cudaGetLastError());
  return LFM_OK;
}

extern "C" int lfm_batched_nlml_grad_unc(lfm_stream_t stream, int64_t B, int64_t N, int G, const double* X,
                                         const double* y, const double* theta_unc, double jitter,
                                         double* out_val, double* out_grad, int* info) {
  if (!out_val || !out_grad) return LFM_ERR_INVALID;
  BatchedArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = (int)N; a.G = G; a.X = X; a.y = y;
  a.u_io = const_cast<double*>(theta_unc);  // read-only in eval mode
  a.jitter = jitter; a.steps = 1; a.total_steps = 1; a.steps_per_epoch = 1;
  a.eval_val = out_val; a.eval_grad = out_grad; a.info = info;
  if (N > 128) return LFM_ERR_UNSUPPORTED;
  return batched_launch((cudaStream_t)stream, a);
}

extern "C" int lfm_batched_fit(lfm_stream_t stream, int64_t B, int64_t N, int G, const double* X, const double* y,
                               double* theta_unc_io, double* adam_state, double jitter, double lr, double b1,
                               double b2, double eps, int first_step, int steps, int total_steps, int fix_params,
                               int steps_per_epoch, double* out_hist, int64_t ld_hist, double* out_theta,
                               int* info) {
  if (steps < 0 || first_step < 0 || steps_per_epoch <= 0) return LFM_ERR_INVALID;
  if (first_step > 0 && !adam_state) return LFM_ERR_INVALID;
  if (N > 128) return LFM_ERR_UNSUPPORTED;
  BatchedArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.N = (int)N; a.G = G; a.X = X; a.y = y; a.u_io = theta_unc_io; a.adam = adam_state;
  a.jitter = jitter; a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps;
  a.first_step = first_step; a.steps = steps; a.total_steps = total_steps; a.fix_params = fix_params;
  a.steps_per_epoch = steps_per_epoch; a.hist = out_hist; a.ld_hist = ld_hist; a.theta_out = out_theta;
  a.info = info;
  return batched_launch((cudaStream_t)stream, a);
}
