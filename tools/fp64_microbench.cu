// Latency / throughput of the FP64 instructions the leaf and batched kernels depend on (one warp, clock64).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rsq(double x) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__global__ void k(double* out, long long* cyc, double seed) {
  double a = seed + threadIdx.x, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) a = fma(a, b, c);
  t1 = clock64(); cyc[0] = t1 - t0;
  // 8 independent DFMA chains
  double x[8]; for (int i = 0; i < 8; ++i) x[i] = a + i;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fma(x[j], b, c);
  t1 = clock64(); cyc[1] = t1 - t0;
  for (int i = 0; i < 8; ++i) a += x[i];
  // dependent shuffle chain (double = 2 x 32-bit shuffles)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) a = __shfl_sync(0xffffffffu, a, (i * 7) & 31) + 0.0;
  t1 = clock64(); cyc[2] = t1 - t0;
  // dependent MUFU.RSQ64H chain
  a = fabs(a) + 1.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) a = rsq(a) + 1.0;
  t1 = clock64(); cyc[3] = t1 - t0;
  // dependent DMMA chain
  double c0 = a, c1 = a;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(b), "d"(c));
  t1 = clock64(); cyc[4] = t1 - t0;
  // 8 independent DMMA accumulators
  double d0[8], d1[8]; for (int i = 0; i < 8; ++i) { d0[i] = c0 + i; d1[i] = c1; }
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0[j]), "+d"(d1[j]) : "d"(b), "d"(c));
  t1 = clock64(); cyc[5] = t1 - t0;
  for (int i = 0; i < 8; ++i) a += d0[i] + d1[i];
  // smem store -> syncwarp -> broadcast load round trip, dependent
  __shared__ double sm[64];
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { sm[threadIdx.x & 31] = a; __syncwarp(); a = sm[(i * 5) & 31] + 1.0; __syncwarp(); }
  t1 = clock64(); cyc[6] = t1 - t0;
  out[threadIdx.x] = a + c0 + c1;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 16); long long h[8];
  for (int warps = 1; warps <= 4; warps *= 2) {   // warps on the same SM (one per scheduler up to 4)
    k<<<1, 32 * warps>>>(out, cyc, 1.0); k<<<1, 32 * warps>>>(out, cyc, 1.0);
    cudaDeviceSynchronize(); cudaMemcpy(h, cyc, 8 * 7, cudaMemcpyDeviceToHost);
    printf("warps %d: DFMA dep %.1f clk | DFMA 8-indep %.2f clk/instr | SHFL.f64 dep %.1f | RSQ64H+DADD dep %.1f | DMMA dep %.1f | DMMA 8-indep %.2f clk/instr | STS+sync+LDS+DADD %.1f\n",
           warps, h[0] / 256.0, h[1] / 512.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 256.0, h[6] / 64.0);
  }
  return 0;
}
