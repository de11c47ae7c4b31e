// Batched small-N path, second kernel: ONE WARP per LFM.
//
// Same mathematics as lfm_batched_kernel (batched.cu: duplicate-row compression to U unique rows with
// multiplicity R, M = c I + R K_u, everything resident in shared memory for all optimiser steps of a
// launch), restructured around three observations from the ncu profile of the CTA-per-LFM kernel
// (profiles/batched_r1_*.md): 35 % of its warp stalls were CTA barriers of phases with fewer than 128
// useful threads, 17 % were instruction-cache misses of the exp/erf code inlined at every pair
// evaluation, and the pair terms were evaluated twice per step (value pass and gradient pass).
//   * one warp owns one problem: every phase is warp-synchronous (shuffles and __syncwarp, no CTA
//     barrier), and ~10 independent problems per SM hide each other's latencies;
//   * the exp/erf factors of h (src/model.py:315-365) are tabulated once per step over
//     (gene, time index, time index) -- the time grid of sim_math.cuh/grid.cu held in shared memory:
//     G T^2 evaluations instead of 2 U^2 (p53: 245 instead of 2520), one erf/erfc call per entry;
//   * rows are mapped to lanes as {lane < U-32, (U-32) + lane}, so the triangular phases touch the
//     U - 32 "extra" rows only while they are short.
// Limits: N <= 128, U <= 40, 3G+2 <= 64, G T^2 <= 2048; anything else runs the CTA-per-LFM kernel.
#include "batched.cuh"


__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void wpair_decode(int p, int& r, int& c) {
  r = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
  while ((r + 1) * (r + 2) / 2 <= p) ++r;
  while (r * (r + 1) / 2 > p) --r;
  c = p - r * (r + 1) / 2;
}
// 1/sqrt(x) for x > 0 normal: MUFU.RSQ64H seed + one third-order step (branch free)
__device__ __forceinline__ double w_rsqrt(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = fma(-(y0 * y0), x, 1.0);
  const double t = fma(e, 0.375, 0.5);
  return fma(t, y0 * e, y0);
}

#define WEX 8             // most "extra" rows beyond 32 the warp kernel takes (U <= 40)
#define XLD (WEX + 1)
struct WarpLayout {
  int ld;
  size_t S, tA1R1, tA1, tG1, g2, inv, utime, e2, c2, q, w, beta, kb, wdiag, sdiag, dsum, th, u, gr, am, av, mu, ys, ring, Msm, Xs, xdiag;
  size_t pts;       // byte offset
  size_t ints;      // byte offset: umap[N], urow[MU], rows_of[N], mflag[N]
  size_t bytes;
};
__host__ __device__ inline WarpLayout warp_layout(int N, int G, int MU, int MT) {
  WarpLayout L;
  const int P = 3 * G + 2;
  L.ld = MU | 1;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 1) & ~(size_t)1; return r; };
  L.S = take((size_t)MU * L.ld > 32 * 33 ? (size_t)MU * L.ld : 32 * 33);  // also holds W22 as Wb[32][33]
  const size_t tab = (size_t)G * MT * MT;
  L.tA1R1 = take(tab); L.tA1 = take(tab); L.tG1 = take(tab);
  L.g2 = take((size_t)G * MT); L.inv = take((size_t)G * G); L.utime = take(MT);
  L.e2 = take((size_t)G * MT); L.c2 = take((size_t)G * MT);
  L.q = take(MU); L.w = take(MU); L.beta = take(MU); L.kb = take(MU); L.wdiag = take(MU); L.sdiag = take(MU);
  L.dsum = take(MU);
  L.th = take(P); L.u = take(P); L.gr = take(P); L.am = take(P); L.av = take(P); L.mu = take(G);
  L.ys = take(N);
  L.ring = take(4 * 32);
  L.Msm = take(WEX * 32);
  L.Xs = take(WEX * XLD);
  L.xdiag = take(WEX);
  L.pts = o * 8;
  size_t b = L.pts + (size_t)MU * sizeof(LfmPoint);
  b = (b + 15) & ~(size_t)15;
  L.ints = b;
  b += sizeof(int) * ((size_t)3 * N + MU);
  b += sizeof(unsigned short) * ((size_t)MU * (MU + 1) / 2 + 2);  // pair table
  L.bytes = (b + 15) & ~(size_t)15;
  return L;
}

__global__ void __launch_bounds__(32) lfm_batched_warp_kernel(BatchedArgs a, int MT) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.N, G = a.G, P = 3 * G + 2;
  const int lane = threadIdx.x;
  const int64_t bidx = blockIdx.x;
  const int MU = a.max_unique;
  const WarpLayout L = warp_layout(N, G, MU, MT);
  double* base = reinterpret_cast<double*>(smem_raw);
  double* S = base + L.S;
  double* tA1R1 = base + L.tA1R1; double* tA1 = base + L.tA1; double* tG1 = base + L.tG1;
  double* g2 = base + L.g2; double* inv = base + L.inv; double* utime = base + L.utime;
  double* e2 = base + L.e2; double* c2 = base + L.c2;
  double* q = base + L.q; double* w = base + L.w; double* beta = base + L.beta; double* kb = base + L.kb;
  double* wdiag = base + L.wdiag; double* sdiag = base + L.sdiag; double* dsum = base + L.dsum;
  double* th = base + L.th; double* u = base + L.u; double* gr = base + L.gr; double* am = base + L.am;
  double* av = base + L.av; double* mu = base + L.mu; double* ys = base + L.ys; double* ring = base + L.ring;
  double* Msm = base + L.Msm; double* Xs = base + L.Xs; double* xdiag = base + L.xdiag;
  LfmPoint* pts = reinterpret_cast<LfmPoint*>(smem_raw + L.pts);
  int* umap = reinterpret_cast<int*>(smem_raw + L.ints);  // row -> unique index       (N)
  int* urow = umap + N;                                   // unique index -> first row (MU)
  int* rows_of = urow + MU;                               // rows of class u: rows_of[u * R + r] (N)
  int* mflag = rows_of + N;                               // 2 * positional block + flag (N)
  unsigned short* pairs = reinterpret_cast<unsigned short*>(mflag + N);  // lower-triangle pair p -> (r << 8) | c

  for (int p = lane; p < P; p += 32) {
    u[p] = a.u_io[bidx * P + p];
    const bool have = a.adam != nullptr && a.first_step > 0;
    am[p] = have ? a.adam[bidx * 2 * P + p] : 0.0;
    av[p] = have ? a.adam[bidx * 2 * P + P + p] : 0.0;
  }
  int fail = 0;
  const int blk = N / G;  // rows per positional mean block (model.py:145)
  // ---- once per launch: duplicate rows (class representative = first identical row), multiplicity ---------
  for (int i = lane; i < N; i += 32) {
    int rep = i;
    const double t0 = a.X[3 * i], g0 = a.X[3 * i + 1], f0 = a.X[3 * i + 2];
    for (int j = 0; j < i; ++j)
      if (a.X[3 * j] == t0 && a.X[3 * j + 1] == g0 && a.X[3 * j + 2] == f0) { rep = j; break; }
    umap[i] = rep;
    ys[i] = a.y[i];
    int m = i / blk;
    if (m > G - 1) m = G - 1;
    mflag[i] = 2 * m + (((int)f0) != 0 ? 1 : 0);
  }
  __syncwarp();
  int U = 0, R = 0, uniform = 1;
  {
    int nrep = 0, cnt0 = 0;
    for (int i = lane; i < N; i += 32) {
      const int rep = umap[i];
      if (rep == i) ++nrep;
      int cnt = 0;
      for (int j = 0; j < N; ++j) cnt += (umap[j] == rep);
      if (i == 0) cnt0 = cnt;
      rows_of[i] = cnt;  // temporarily the multiplicity of row i's class
    }
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) nrep += __shfl_xor_sync(0xffffffffu, nrep, o);
    cnt0 = __shfl_sync(0xffffffffu, cnt0, 0);
    int bad = 0;
    for (int i = lane; i < N; i += 32) bad |= (rows_of[i] != cnt0);
    bad = __any_sync(0xffffffffu, bad);
    U = nrep; R = cnt0; uniform = !bad;
    if (!uniform || R == 1) { U = N; R = 1; }
  }
  __syncwarp();
  if (U > MU) { U = 0; fail = -1; }  // caller's unique-row bound was wrong: refuse (info = -1)
  if (fail == 0) {
    // compact index of every class (ordered by representative row), its rows in ascending order
    for (int i = lane; i < N; i += 32) {
      int idx = i;
      if (R > 1) {
        const int rep = umap[i];
        idx = 0;
        for (int j = 0; j < rep; ++j) idx += (umap[j] == j);
      }
      mflag[i] |= idx << 8;  // park the compact index above the 8 low bits (2 * m + flag < 256)
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) {
      const int idx = mflag[i] >> 8;
      int ord = 0;
      if (R > 1) for (int j = 0; j < i; ++j) ord += ((mflag[j] >> 8) == idx);
      rows_of[idx * R + ord] = i;
      if (ord == 0) urow[idx] = i;
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) { umap[i] = mflag[i] >> 8; mflag[i] &= 255; }
    __syncwarp();
  }
  // ---- distinct times of the unique rows -> pts[].ti, utime[] ------------------------------------------------
  int Tu = 0;
  if (fail == 0) {
    int nfirst = 0;
    for (int r = lane; r < U; r += 32) {
      const double t = a.X[3 * urow[r]];
      int first = 1;
      for (int j = 0; j < r; ++j) if (a.X[3 * urow[j]] == t) { first = 0; break; }
      nfirst += first;
      pts[r].flag = first;  // temporary marker
    }
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) nfirst += __shfl_xor_sync(0xffffffffu, nfirst, o);
    Tu = nfirst;
    if (Tu > MT) { fail = -2; }  // caller's time-grid bound was wrong: refuse (info = -2)
    else {
      for (int r = lane; r < U; r += 32) {
        const double t = a.X[3 * urow[r]];
        int idx = 0, rep = r;
        for (int j = 0; j < r; ++j) if (a.X[3 * urow[j]] == t) { rep = j; break; }
        for (int j = 0; j < rep; ++j) idx += pts[j].flag;
        pts[r].ti = idx;
        if (rep == r) utime[idx] = t;
      }
    }
    __syncwarp();
  }
  if (fail != 0) { U = 0; Tu = 0; }

  const int ld = U | 1;
  const int npairs = U * (U + 1) / 2;
  for (int p = lane; p < npairs; p += 32) {
    int r, cc;
    wpair_decode(p, r, cc);
    pairs[p] = (unsigned short)((r << 8) | cc);
  }
  __syncwarp();
  const bool eval_only = a.eval_val != nullptr;
  const int nsteps = eval_only ? 1 : a.steps;
  const double dR = (double)R;
  const int ex = U > 32 ? U - 32 : 0;  // rows [0, ex) are the "extra" rows: lane -> rows {lane < ex, ex + lane}
  const int TT = Tu * Tu;

  for (int sidx = 0; sidx < nsteps; ++sidx) {
    const int step = a.first_step + sidx;
    // ---- A. constrain -----------------------------------------------------------------------
    for (int p = lane; p < P; p += 32) th[p] = (p == 3 * G) ? lfm_l_forward(u[p]) : lfm_softplus(u[p]);
    __syncwarp();
    const double l = th[3 * G], inv_l = 1.0 / l, sigma = th[3 * G + 1];
    const double c = a.jitter + sigma * sigma;
    // ---- B. unique points, q = P^T z, z^T z ------------------------------------------------------
    for (int m = lane; m < G; m += 32) mu[m] = th[2 * G + m] / th[m];
    for (int r = lane; r < U; r += 32) {
      const int ti = pts[r].ti;
      LfmPoint p = lfm_make_point(a.X + 3 * urow[r], G, th, th + G, l, true);
      p.ti = ti;
      pts[r] = p;
    }
    __syncwarp();
    double zz = 0.0;
    for (int i = lane; i < N; i += 32) {
      const double zi = ys[i] - mu[mflag[i] >> 1] * (double)(mflag[i] & 1);
      zz += zi * zi;
    }
    zz = wsum(zz);
    for (int r = lane; r < U; r += 32) {
      double acc = 0.0;
      for (int k = 0; k < R; ++k) {
        const int i = rows_of[r * R + k];
        acc += ys[i] - mu[mflag[i] >> 1] * (double)(mflag[i] & 1);
      }
      q[r] = acc;
    }
    // ---- C. time-grid tables: per (gene b, time ia) the x2 terms, per (b, ia, ib) the pair terms -------------
    for (int e = lane; e < G * Tu; e += 32) {
      const int b = e / Tu, ia = e % Tu;
      const double gam = th[b] * l * 0.5;
      const double x2 = utime[ia] * inv_l + gam;
      e2[b * MT + ia] = erf(x2);
      c2[b * MT + ia] = erfc(fabs(x2));
      g2[b * MT + ia] = __dmul_rn(LFM_TWO_OVER_SQRT_PI, exp(-x2 * x2));
    }
    for (int e = lane; e < G * G; e += 32) inv[e] = 1.0 / (th[e / G] + th[e % G]);
    __syncwarp();
    for (int e = lane; e < G * TT; e += 32) {
      const int b = e / TT, r2 = e % TT, ia = r2 / Tu, ib = r2 % Tu;
      const double d_b = th[b];
      const double gam = d_b * l * 0.5;
      const double ta = utime[ia], tb = utime[ib];
      const double delta = tb - ta;
      const double A1 = exp(-d_b * delta);
      const double x1 = delta * inv_l - gam;
      const double x2 = ta * inv_l + gam;
      // lfm_erfsum(x1, x2) with the x2 half read from the (b, ia) tables: identical values, half the calls
      double R1;
      if (x1 * x2 < 0.0 && fmin(fabs(x1), fabs(x2)) > 0.5) {
        const double cx1 = erfc(fabs(x1));
        R1 = (x1 < x2) ? (cx1 - c2[b * MT + ia]) : (c2[b * MT + ia] - cx1);  // erfc(-n) - erfc(p)
      } else {
        R1 = erf(x1) + e2[b * MT + ia];
      }
      const size_t o = (size_t)b * MT * MT + (size_t)ia * MT + ib;
      tA1[o] = A1;
      tA1R1[o] = __dmul_rn(A1, R1);
      tG1[o] = __dmul_rn(LFM_TWO_OVER_SQRT_PI, exp(-x1 * x1));
    }
    __syncwarp();
    // h(pa, pb) from the shared-memory tables
    auto pair_terms = [&](const LfmPoint& pa, const LfmPoint& pb) {
      LfmPairTerms pt;
      const size_t o = (size_t)pb.gene * MT * MT + (size_t)pa.ti * MT + pb.ti;
      pt.A1 = tA1[o]; pt.A1R1 = tA1R1[o]; pt.g1 = tG1[o];
      pt.g2 = g2[pb.gene * MT + pa.ti];
      pt.inv = inv[pa.gene * G + pb.gene];
      return pt;
    };
    // ---- D. M = c I + R K_u (lower + diagonal) -------------------------------------------------------
    for (int p = lane; p < npairs; p += 32) {
      const int r = pairs[p] >> 8, cc = pairs[p] & 255;
      const LfmPoint pi = pts[r], pj = pts[cc];
      double H1, H2, u0, u1, u2;
      lfm_h_core<false>(pj, pi, l, inv_l, pair_terms(pj, pi), H1, u0, u1, u2);
      lfm_h_core<false>(pi, pj, l, inv_l, pair_terms(pi, pj), H2, u0, u1, u2);
      double k = dR * (pi.s * pj.s * (LFM_SQRT_PI * 0.5 * l) * (H1 + H2));
      if (r == cc) k += c;
      S[r * ld + cc] = k;
    }
    __syncwarp();
    // ---- E. M^-1 and log det M.  M = [[M11, M21^T], [M21, M22]] with M22 the trailing n2 = U - ex (<= 32) rows.
    // M22 is factorised AND inverted in registers (lane i owns row i of M22 / L22 and column i of W22 = L22^-1; the
    // pivot loop is fully unrolled; lanes >= n2 carry identity rows); the ex <= 8 leading rows enter through the
    // Schur complement, every per-lane quantity in registers with compile-time indices:
    //   B = M22^-1 M21,  S11 = M11 - M21^T B,  X11 = S11^-1,  Y = B X11,
    //   M^-1 = [[X11, -Y^T], [-Y, M22^-1 + Y B^T]],   log det M = log det M22 + log det S11.
    double logdetM;
    {
      const int n2 = U - ex;
      double am[32], yw[32], m21[WEX];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        am[j] = (lane < n2) ? ((j <= lane) ? S[(ex + lane) * ld + ex + j] : 0.0) : ((j == lane) ? 1.0 : 0.0);
        yw[j] = (j == lane) ? 1.0 : 0.0;
      }
#pragma unroll
      for (int j = 0; j < WEX; ++j) {
        m21[j] = (j < ex && lane < n2) ? S[(ex + lane) * ld + j] : 0.0;
        Msm[j * 32 + lane] = m21[j];                       // M21 transposed: Msm[j][c] = M21[c][j]
      }
      if (lane < ex)
        for (int b1 = 0; b1 <= lane; ++b1) Xs[lane * XLD + b1] = S[lane * ld + b1];
      __syncwarp();  // S is free from here on: it becomes Wb[32][33]
      double my_rk = 1.0;
      double akk = __shfl_sync(0xffffffffu, am[0], 0);
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        if (!(akk > 0.0) && fail == 0) fail = ex + k + 1;
        const double rk = w_rsqrt(akk);
        const double lik = (lane == k) ? akk * rk : am[k] * rk;
        double akk_next = 0.0;
        if (k + 1 < 32) akk_next = __shfl_sync(0xffffffffu, fma(-lik, lik, am[k + 1]), k + 1);
        if (lane == k) my_rk = rk;
        am[k] = lik;
        double* Tk = ring + (k & 3) * 32;
        Tk[lane] = (lane >= k) ? lik : 0.0;
        const double wk = yw[k] * rk;
        yw[k] = wk;
        __syncwarp();
        if (k + 1 < 32) {
          if ((k + 1) & 1) {
            const double l1 = Tk[k + 1];
            am[k + 1] = fma(-lik, l1, am[k + 1]);
            yw[k + 1] = fma(-l1, wk, yw[k + 1]);
          }
#pragma unroll
          for (int j = (k + 2) & ~1; j < 32; j += 2) {
            const double2 l2 = *reinterpret_cast<const double2*>(Tk + j);
            am[j] = fma(-lik, l2.x, am[j]);
            am[j + 1] = fma(-lik, l2.y, am[j + 1]);
            yw[j] = fma(-l2.x, wk, yw[j]);
            yw[j + 1] = fma(-l2.y, wk, yw[j + 1]);
          }
        }
        akk = akk_next;
      }
      double ld22 = (lane < n2) ? -log(my_rk) : 0.0;   // log L_kk = -log(1 / L_kk)
      ld22 = 2.0 * wsum(ld22);
      // W22 -> Wb[i][c] (constant stride 33; padded lanes carry exact zeros / identity)
      double* Wb = S;
#pragma unroll
      for (int i = 0; i < 32; ++i) Wb[i * 33 + lane] = yw[i];
      __syncwarp();
      // row `lane` of M22^-1 = W22^T W22 (complete row, both sides of the diagonal)
      double mrow[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        double s0 = 0.0;
#pragma unroll
        for (int k = c; k < 32; ++k) s0 = fma(yw[k], Wb[k * 33 + c], s0);
        mrow[c] = s0;
      }
      double ld11 = 0.0;
      double yrow[WEX];
#pragma unroll
      for (int j = 0; j < WEX; ++j) yrow[j] = 0.0;
      if (ex > 0) {
        double bcol[WEX];
#pragma unroll
        for (int j = 0; j < WEX; ++j) {
          double s0 = 0.0;
          if (j < ex) {
#pragma unroll
            for (int c = 0; c < 32; ++c) s0 = fma(mrow[c], Msm[j * 32 + c], s0);
          }
          bcol[j] = s0;   // B[lane][j]
        }
        // S11 = M11 - M21^T B (lower), by lane 0 into Xs
#pragma unroll
        for (int a1 = 0; a1 < WEX; ++a1)
#pragma unroll
          for (int b1 = 0; b1 <= a1; ++b1)
            if (a1 < ex) {
              const double t = wsum(m21[a1] * bcol[b1]);
              if (lane == 0) Xs[a1 * XLD + b1] -= t;
            }
        __syncwarp();
        // S11 = L11 L11^T (left-looking, lane per row); W11 = L11^-1 in the upper triangle; X11 = W11^T W11
        for (int k = 0; k < ex; ++k) {
          double v = 0.0;
          const bool on = lane >= k && lane < ex;
          if (on) {
            double s0 = 0.0;
            for (int m = 0; m < k; ++m) s0 = fma(Xs[lane * XLD + m], Xs[k * XLD + m], s0);
            v = Xs[lane * XLD + k] - s0;
          }
          const double piv = __shfl_sync(0xffffffffu, v, k);
          if (!(piv > 0.0) && fail == 0) fail = k + 1;
          const double rkk = w_rsqrt(piv);
          if (on) Xs[lane * XLD + k] = (lane == k) ? piv * rkk : v * rkk;
          if (lane == 0) { xdiag[k] = rkk; ld11 -= log(rkk); }
          __syncwarp();
        }
        ld11 = 2.0 * __shfl_sync(0xffffffffu, ld11, 0);
        if (lane < ex) {  // column `lane` of W11 into row `lane` of the upper triangle: Xs[c][i] = W11[i][c], i > c
          const int cc = lane;
          for (int i = cc + 1; i < ex; ++i) {
            double s0 = Xs[i * XLD + cc] * xdiag[cc];
            for (int k = cc + 1; k < i; ++k) s0 = fma(Xs[i * XLD + k], Xs[cc * XLD + k], s0);
            Xs[cc * XLD + i] = -s0 * xdiag[i];
          }
        }
        __syncwarp();
        if (lane < ex) {  // row `lane` of X11 = W11^T W11 (lower + diagonal) -> registers, then in place
          const int a1 = lane;
          double xr[WEX];
#pragma unroll
          for (int b1 = 0; b1 < WEX; ++b1) {
            double s0 = 0.0;
            if (b1 <= a1) {
              s0 = xdiag[a1] * ((a1 == b1) ? xdiag[a1] : Xs[b1 * XLD + a1]);
              for (int k = a1 + 1; k < ex; ++k) s0 = fma(Xs[a1 * XLD + k], Xs[b1 * XLD + k], s0);
            }
            xr[b1] = s0;
          }
#pragma unroll
          for (int b1 = 0; b1 < WEX; ++b1)
            if (b1 <= a1) Xs[a1 * XLD + b1] = xr[b1];   // lower incl. diagonal; the upper triangle (W11) stays readable
        }
        __syncwarp();
        // Y = B X11 (X11 symmetric, read from the lower triangle), B -> Msm for the rank-ex correction
#pragma unroll
        for (int a1 = 0; a1 < WEX; ++a1) {
          double s0 = 0.0;
          if (a1 < ex) {
#pragma unroll
            for (int b1 = 0; b1 < WEX; ++b1)
              if (b1 < ex) s0 = fma(bcol[b1], (b1 <= a1) ? Xs[a1 * XLD + b1] : Xs[b1 * XLD + a1], s0);
          }
          yrow[a1] = s0;   // Y[lane][a1]
        }
#pragma unroll
        for (int j = 0; j < WEX; ++j) Msm[j * 32 + lane] = bcol[j];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          double s0 = mrow[c];
#pragma unroll
          for (int a1 = 0; a1 < WEX; ++a1)
            if (a1 < ex) s0 = fma(yrow[a1], Msm[a1 * 32 + c], s0);   // + Y[lane][a] B[c][a]
          mrow[c] = s0;
        }
      }
      __syncwarp();  // every lane is done with Wb: S receives M^-1 (lower triangle, runtime ld) and sdiag
      if (lane < n2) {
        double* row = S + (ex + lane) * ld;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c < lane) row[ex + c] = mrow[c];
          else if (c == lane) sdiag[ex + lane] = mrow[c];
        }
#pragma unroll
        for (int a1 = 0; a1 < WEX; ++a1)
          if (a1 < ex) row[a1] = -yrow[a1];               // M^-1_21 = -Y
      }
      if (lane < ex) {
        for (int b1 = 0; b1 < lane; ++b1) S[lane * ld + b1] = Xs[lane * XLD + b1];
        sdiag[lane] = Xs[lane * XLD + lane];
      }
      __syncwarp();
      logdetM = ld22 + ld11;
    }
    // ---- F. beta = M^-1 q (M^-1 symmetric: lower triangle + sdiag), K_u beta = (q - c beta) / R ---------------
    double qkb = 0.0, kbkb = 0.0;
    for (int r = lane; r < U; r += 32) {
      double acc = sdiag[r] * q[r];
      const double* rowr = S + r * ld;
      for (int cc = 0; cc < r; ++cc) acc = fma(rowr[cc], q[cc], acc);
      for (int cc = r + 1; cc < U; ++cc) acc = fma(S[cc * ld + r], q[cc], acc);
      beta[r] = acc;
      const double kbv = (q[r] - c * acc) / dR;
      kb[r] = kbv;
      qkb += q[r] * kbv;
      kbkb += kbv * kbv;
    }
    qkb = wsum(qkb);
    kbkb = wsum(kbkb);
    const double quad = (zz - qkb) / c;
    const double nlml = 0.5 * ((double)N * LFM_LOG_2PI + (double)(N - U) * log(c) + logdetM + quad);
    __syncwarp();
    // ---- I. fused derivative contraction over the lower triangle of the unique pairs --------------------
    double dl_part = 0.0;
    for (int r = lane; r < U; r += 32) dsum[r] = 0.0;
    __syncwarp();
    for (int p = lane; p < npairs; p += 32) {
      const int r = pairs[p] >> 8, cc = pairs[p] & 255;
      const double minv = (r == cc) ? sdiag[r] : S[r * ld + cc];
      const double wgt = ((r == cc) ? 0.5 : 1.0) * (dR * minv - beta[r] * beta[cc]);
      const LfmPoint pi = pts[r], pj = pts[cc];
      double H1, dH1_da, dH1_db, dH1_dl, H2, dH2_da, dH2_db, dH2_dl;
      lfm_h_core<true>(pj, pi, l, inv_l, pair_terms(pj, pi), H1, dH1_da, dH1_db, dH1_dl);
      lfm_h_core<true>(pi, pj, l, inv_l, pair_terms(pi, pj), H2, dH2_da, dH2_db, dH2_dl);
      const double mult = pi.s * pj.s * (LFM_SQRT_PI * 0.5 * l);
      const double k = mult * (H1 + H2);
      const double dr = mult * (dH1_db + dH2_da);
      const double dc = mult * (dH1_da + dH2_db);
      const double dl = mult * (dH1_dl + dH2_dl) + k * inv_l;
      dl_part += wgt * dl;
      if (r == cc) dsum[r] = wgt * (dr + dc);
      else { S[r * ld + cc] = wgt * dr; S[cc * ld + r] = wgt * dc; }
    }
    const double gl = wsum(dl_part);
    __syncwarp();
    for (int r = lane; r < U; r += 32) {  // per-point totals: full row sums, fixed order
      const double* rp = S + r * ld;
      double s0 = dsum[r], s1 = 0.0;
      int cc = 0;
      for (; cc + 2 <= U; cc += 2) {
        s0 += (cc == r) ? 0.0 : rp[cc];
        s1 += (cc + 1 == r) ? 0.0 : rp[cc + 1];
      }
      if (cc < U) s0 += (cc == r) ? 0.0 : rp[cc];
      dsum[r] = s0 + s1;
    }
    __syncwarp();
    // ---- J. fold by gene; mean-function terms; sigma ---------------------------------------------
    for (int m = 0; m < G; ++m) {
      double gd = 0.0, gs = 0.0, asum = 0.0;
      for (int i = lane; i < U; i += 32) {
        if (pts[i].gene == m) {
          gd += dsum[i];
          gs += 1.0 - c * sdiag[i] - beta[i] * kb[i];
        }
      }
      // asum_m = sum_{i in positional block m} alpha_i,  alpha_i = (z_i - (K_u beta)_{u(i)}) / c
      for (int i = m * blk + lane; i < (m + 1) * blk; i += 32)
        asum += ys[i] - mu[m] * (double)(mflag[i] & 1) - kb[umap[i]];
      gd = wsum(gd); gs = wsum(gs); asum = wsum(asum) / c;
      if (lane == 0) {
        const double D = th[m], Sm = th[G + m], Bm = th[2 * G + m];
        gr[m] = gd + asum * Bm / (D * D);
        gr[G + m] = gs / Sm;
        gr[2 * G + m] = -asum / D;
      }
    }
    {
      double tr = 0.0;
      for (int i = lane; i < U; i += 32) tr += sdiag[i];
      tr = wsum(tr);
      if (lane == 0) {
        const double trSinv = ((double)(N - U) + c * tr) / c;
        const double aa = (zz - 2.0 * qkb + dR * kbkb) / (c * c);
        gr[3 * G] = gl;
        gr[3 * G + 1] = sigma * (trSinv - aa);
      }
    }
    __syncwarp();
    // ---- K. chain rule, Adam, hook ----------------------------------------------------------------
    const bool bad = fail != 0;
    for (int p = lane; p < P; p += 32) {
      const double sg = lfm_sigmoid(u[p]);
      const double jac = (p == 3 * G) ? (LFM_L_HIGH - LFM_L_LOW) * sg * (1.0 - sg) : sg;
      double g = gr[p] * jac;
      if (bad) g = nan("");
      if (eval_only) {
        a.eval_grad[bidx * P + p] = g;
      } else {
        const double m1 = a.b1 * am[p] + (1.0 - a.b1) * g;
        const double v1 = a.b2 * av[p] + (1.0 - a.b2) * g * g;
        am[p] = m1; av[p] = v1;
        const double mhat = m1 / (1.0 - pow(a.b1, (double)(step + 1)));
        const double vhat = v1 / (1.0 - pow(a.b2, (double)(step + 1)));
        double un = u[p] - a.lr * mhat / (sqrt(vhat) + a.eps);
        if (a.fix_params && (step % a.steps_per_epoch) == 0 && G > 3) {
          if (p == G + 3) un = 1.0;  // true_s[3]  (trainer.py:152, unconstrained space: SURVEY Q5)
          if (p == 3) un = 0.8;      // true_d[3]  (trainer.py:153)
        }
        u[p] = un;
      }
    }
    if (lane == 0) {
      const double v = bad ? nan("") : nlml;
      if (eval_only) a.eval_val[bidx] = v;
      else if (a.hist) a.hist[bidx * a.ld_hist + step] = v;
    }
    __syncwarp();
  }

  if (!eval_only) {
    for (int p = lane; p < P; p += 32) {
      a.u_io[bidx * P + p] = u[p];
      if (a.adam) { a.adam[bidx * 2 * P + p] = am[p]; a.adam[bidx * 2 * P + P + p] = av[p]; }
      if (a.theta_out && a.first_step + a.steps >= a.total_steps) {
        double t = (p == 3 * G) ? lfm_l_forward(u[p]) : lfm_softplus(u[p]);  // trainer.py:218
        if (a.fix_params && G > 3) {                                         // trainer.py:219-220
          if (p == G + 3) t = 1.0;
          if (p == 3) t = 0.8;
        }
        a.theta_out[bidx * P + p] = t;
      }
    }
  }
  if (lane == 0 && a.info) {
    if (a.first_step == 0 || eval_only) a.info[bidx] = fail;
    else if (fail) a.info[bidx] = fail;
  }
  if (lane == 0 && a.best_key && !eval_only && a.hist && a.steps > 0) {
    const double v = a.hist[bidx * a.ld_hist + a.first_step + a.steps - 1];
    if (v == v) atomicMin(a.best_key, lfm_loss_key(v));
  }
}

// Launch the warp-per-LFM kernel if the problem fits its limits; returns LFM_ERR_UNSUPPORTED otherwise
// (the caller then runs the CTA-per-LFM kernel).
int lfm_batched_warp_launch(cudaStream_t st, const BatchedArgs& a, int time_grid) {
  const int P = 3 * a.G + 2;
  const int MU = a.max_unique;
  if (time_grid <= 0 || MU <= 0 || MU > 32 + WEX || a.N > 128 || P > 64) return LFM_ERR_UNSUPPORTED;
  if ((long long)a.G * time_grid * time_grid > 2048 || a.G > 127) return LFM_ERR_UNSUPPORTED;
  const WarpLayout L = warp_layout(a.N, a.G, MU, time_grid);
  if (L.bytes > 100 * 1024) return LFM_ERR_UNSUPPORTED;
  static size_t conf = 0;
  if (L.bytes > conf) {
    LFM_CUDA_OK(cudaFuncSetAttribute(lfm_batched_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes));
    conf = L.bytes;
  }
  lfm_batched_warp_kernel<<<(unsigned)a.B, 32, L.bytes, st>>>(a, time_grid);
  LFM_LAUNCHED(1);
  LFM_CUDA_OK(cudaGetLastError());
  return LFM_OK;
}
