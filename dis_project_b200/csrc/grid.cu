// Distinct-time tables for gridded data sets ("time grid").
//
// The reference evaluates the SIM kernel independently for every pair of rows (vmap(vmap(kernel)),
// src/model.py:372-394).  Its data layout, however, observes every gene on the same few time points
// (dataset.py:380-391), and the expensive factors of h (model.py:315-365) -- exp(-D_k (t' - t)),
// erf((t' - t)/l - gamma_k) + erf(t/l + gamma_k) and the Gaussians of the same arguments that the
// gradient needs -- depend on (gene k, t, t') only.  This file
//   1. finds the distinct times of X on the device (open-addressing hash on the bit pattern of t),
//      giving every row its time index;
//   2. tabulates the pair terms once per evaluation: G T^2 entries instead of N^2 = (G T)^2.
// The tile kernels of gram.cu read the tables when the number of distinct times fits the caller's bound
// (`time_grid`), and fall back to direct evaluation otherwise -- decided on the device, no host sync.
// The table entries are produced by the same lfm_pair_terms() the direct path calls.
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "sim_math.cuh"

#define GRID_EMPTY 0xFFFFFFFFFFFFFFFFull

struct LfmGridWs {
  unsigned long long* slots;  // [H]
  int* slot_idx;              // [H]
  int* slot_of;               // [N]
  int* tidx;                  // [N]
  int* count;                 // [1]
  double* utime;              // [Tu]
  double *A1R1, *A1R1t, *Qd, *Qdt, *Rl, *Rlt, *inv;
  int64_t H;
  size_t total_doubles;
};

static int64_t grid_hash_size(int64_t N) {
  int64_t h = 1024;
  while (h < 4 * N) h <<= 1;
  return h;
}

static LfmGridWs grid_layout(int64_t N, int G, int64_t Tu, void* base) {
  LfmGridWs s;
  s.H = grid_hash_size(N);
  double* p = reinterpret_cast<double*>(base);
  size_t off = 0;
  auto take = [&](size_t n) { double* r = p ? p + off : nullptr; off += (n + 1) & ~(size_t)1; return r; };
  s.slots = reinterpret_cast<unsigned long long*>(take((size_t)s.H));
  s.slot_idx = reinterpret_cast<int*>(take((size_t)s.H / 2 + 1));
  s.slot_of = reinterpret_cast<int*>(take((size_t)N / 2 + 1));
  s.tidx = reinterpret_cast<int*>(take((size_t)N / 2 + 1));
  s.count = reinterpret_cast<int*>(take(2));
  s.utime = take((size_t)Tu);
  const size_t tab = (size_t)G * Tu * Tu;
  s.A1R1 = take(tab); s.A1R1t = take(tab);
  s.Qd = take(tab); s.Qdt = take(tab);
  s.Rl = take(tab); s.Rlt = take(tab);
  s.inv = take((size_t)G * G);
  s.total_doubles = off;
  return s;
}

size_t lfm_grid_ws_doubles(int64_t N, int G, int64_t Tu) {
  if (Tu <= 0) return 0;
  return grid_layout(N, G, Tu, nullptr).total_doubles;
}

__device__ __forceinline__ unsigned long long grid_key(double t) {
  if (t != t) return 0x7ff8000000000000ull;  // every NaN is one key (and never GRID_EMPTY)
  if (t == 0.0) t = 0.0;                     // -0 -> +0
  return (unsigned long long)__double_as_longlong(t);
}

__global__ void lfm_grid_insert_kernel(int64_t N, const double* __restrict__ X, unsigned long long* __restrict__ slots,
                                       int64_t H, int* __restrict__ slot_of) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const unsigned long long key = grid_key(X[3 * i]);
  unsigned long long h = key * 0x9E3779B97F4A7C15ull;
  int64_t s = (int64_t)((h >> 20) & (unsigned long long)(H - 1));
  for (;;) {
    const unsigned long long prev = atomicCAS(&slots[s], GRID_EMPTY, key);
    if (prev == GRID_EMPTY || prev == key) break;
    s = (s + 1) & (H - 1);
  }
  slot_of[i] = (int)s;
}
__global__ void lfm_grid_compact_kernel(int64_t H, const unsigned long long* __restrict__ slots,
                                        int* __restrict__ slot_idx, int* __restrict__ count,
                                        double* __restrict__ utime, int Tu) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= H) return;
  const unsigned long long key = slots[s];
  if (key == GRID_EMPTY) return;
  const int idx = atomicAdd(count, 1);
  slot_idx[s] = idx;
  if (idx < Tu) utime[idx] = __longlong_as_double((long long)key);
}
__global__ void lfm_grid_assign_kernel(int64_t N, const int* __restrict__ slot_of, const int* __restrict__ slot_idx,
                                       int* __restrict__ tidx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= N) return;
  tidx[i] = slot_idx[slot_of[i]];
}

// one thread per (gene b, ia, ib)
template <bool GRAD>
__global__ void __launch_bounds__(256) lfm_grid_tables_kernel(LfmGridWs w, int G, int Tu,
                                                            const double* __restrict__ theta) {
  const int cnt = *w.count;
  if (cnt > Tu) return;  // bound exceeded: the tile kernels evaluate directly
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)Tu * Tu;
  if (idx < (int64_t)G * G) {
    const int a = (int)(idx / G), b = (int)(idx % G);
    w.inv[idx] = 1.0 / (theta[a] + theta[b]);
  }
  if (idx >= (int64_t)G * per) return;
  const int b = (int)(idx / per);
  const int r = (int)(idx % per);
  const int ia = r / Tu, ib = r % Tu;
  if (ia >= cnt || ib >= cnt) return;
  const double l = theta[3 * G];
  const double inv_l = 1.0 / l;
  const double d_b = theta[b];
  const double gam = d_b * l * 0.5;
  const LfmPairTerms pt = lfm_pair_terms<GRAD>(w.utime[ia], w.utime[ib], d_b, gam, inv_l);
  const size_t o = (size_t)b * per + (size_t)ia * Tu + ib;
  const size_t ot = (size_t)b * per + (size_t)ib * Tu + ia;
  w.A1R1[o] = pt.A1R1; w.A1R1t[ot] = pt.A1R1;
  if (GRAD) {
    // the (gene b, u, v)-only parts of dH/dd_b and dH/dl (sim_math.cuh: LfmGrid, lfm_h_tab_grad)
    const double u = w.utime[ia], delta = w.utime[ib] - u;
    const double hl = 0.5 * l, hd = 0.5 * d_b, il2 = inv_l * inv_l;
    const double Qd = (gam * l - delta) * pt.A1R1 + pt.A1 * hl * (pt.g2 - pt.g1);
    const double Rl = gam * d_b * pt.A1R1 + pt.A1 * (pt.g1 * (-delta * il2 - hd) + pt.g2 * (-u * il2 + hd));
    w.Qd[o] = Qd; w.Qdt[ot] = Qd;
    w.Rl[o] = Rl; w.Rlt[ot] = Rl;
  }
}

// Build the time index and the tables for this evaluation.  `grid` receives the device view the tile
// kernels take by value.  Tu == 0 disables everything (grid.Tu = 0).
int lfm_grid_build(cudaStream_t st, int64_t N, int G, const double* X, const double* theta, int64_t Tu, bool grad,
                   void* ws, LfmGrid* grid) {
  LfmGrid g;
  memset(&g, 0, sizeof(g));
  g.G = G;
  if (Tu <= 0 || !ws) { *grid = g; return LFM_OK; }
  LfmGridWs w = grid_layout(N, G, Tu, ws);
  LFM_CUDA_OK(cudaMemsetAsync(w.slots, 0xFF, sizeof(unsigned long long) * (size_t)w.H, st));
  LFM_CUDA_OK(cudaMemsetAsync(w.count, 0, 2 * sizeof(int), st));
  lfm_grid_insert_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, X, w.slots, w.H, w.slot_of);
  lfm_grid_compact_kernel<<<(unsigned)((w.H + 255) / 256), 256, 0, st>>>(w.H, w.slots, w.slot_idx, w.count, w.utime,
                                                                        (int)Tu);
  lfm_grid_assign_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, w.slot_of, w.slot_idx, w.tidx);
  int64_t total = (int64_t)G * Tu * Tu;
  if (total < (int64_t)G * G) total = (int64_t)G * G;  // the first G^2 threads also fill the 1/(d_a+d_b) table
  if (grad) lfm_grid_tables_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, G, (int)Tu, theta);
  else lfm_grid_tables_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, G, (int)Tu, theta);
  LFM_LAUNCHED(4);
  LFM_CUDA_OK(cudaGetLastError());
  g.tidx = w.tidx; g.count = w.count; g.Tu = (int)Tu;
  g.A1R1 = w.A1R1; g.A1R1t = w.A1R1t; g.Qd = w.Qd; g.Qdt = w.Qdt; g.Rl = w.Rl; g.Rlt = w.Rlt; g.inv = w.inv;
  *grid = g;
  return LFM_OK;
}

// Host helper: number of distinct times in a HOST copy of X (the `time_grid` bound callers pass).
extern "C" int64_t lfm_count_distinct_times(int64_t N, const double* X_host) {
  if (N <= 0 || !X_host) return 0;
  double* t = (double*)malloc(sizeof(double) * (size_t)N);
  if (!t) return 0;
  // insertion into a sorted prefix: O(N log T) comparisons + O(T^2) moves while T stays small (the useful case);
  // once more than 512 distinct values have been seen, sort the whole column instead (O(N log N), never O(N^2))
  int64_t cnt = 0;
  bool many = false;
  for (int64_t i = 0; i < N && !many; ++i) {
    const double v = X_host[3 * i];
    int64_t lo = 0, hi = cnt;
    while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (t[mid] < v) lo = mid + 1; else hi = mid; }
    if (lo < cnt && t[lo] == v) continue;
    if (cnt == 512) { many = true; break; }
    for (int64_t k = cnt; k > lo; --k) t[k] = t[k - 1];
    t[lo] = v;
    ++cnt;
  }
  if (many) {
    for (int64_t i = 0; i < N; ++i) t[i] = X_host[3 * i];
    std::sort(t, t + N);
    cnt = std::unique(t, t + N) - t;
  }
  free(t);
  return cnt;
}
