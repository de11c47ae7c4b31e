// Probe: SM partitioning with green contexts (driver API through cudaGetDriverEntryPoint), runtime-API launches into
// green-context streams, an 8-CTA cluster in the small partition, stream capture across the partitions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o greenctx_probe tools/greenctx_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>
#define RT(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("RT FAIL %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)
#define DR(x) do { CUresult e_ = (x); if (e_ != CUDA_SUCCESS) { printf("DR FAIL %s line %d: %d\n", #x, __LINE__, (int)e_); return 1; } } while (0)
template <class F> static bool ep(const char* name, F& f) {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point %s\n", name); return false; }
  f = (F)p; return true;
}
__global__ void who(int* out, long long spin) {
  unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) out[blockIdx.x] = (int)smid;
  const long long t0 = clock64();
  while (clock64() - t0 < spin) { }
}
__global__ void __cluster_dims__(8, 1, 1) who_cluster(int* out) {
  extern __shared__ double sm[];
  unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
  sm[threadIdx.x] = 1.0;
  if (threadIdx.x == 0) out[blockIdx.x] = (int)smid;
}
static void show(const char* tag, const std::vector<int>& v) {
  std::set<int> s(v.begin(), v.end());
  printf("%s: %zu CTAs on %zu SMs:", tag, v.size(), s.size());
  int n = 0; for (int x : s) { if (n++ < 40) printf(" %d", x); } printf("\n");
}
int main() {
  RT(cudaSetDevice(0)); RT(cudaFree(0));
  CUresult (*pGetRes)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*pSplit)(CUdevResource*, unsigned*, const CUdevResource*, CUdevResource*, unsigned, unsigned) = nullptr;
  CUresult (*pDesc)(CUdevResourceDesc*, CUdevResource*, unsigned) = nullptr;
  CUresult (*pCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned) = nullptr;
  CUresult (*pStream)(CUstream*, CUgreenCtx, unsigned, int) = nullptr;
  CUresult (*pDevGet)(CUdevice*, int) = nullptr;
  if (!ep("cuDeviceGetDevResource", pGetRes) || !ep("cuDevSmResourceSplitByCount", pSplit) || !ep("cuDevResourceGenerateDesc", pDesc) ||
      !ep("cuGreenCtxCreate", pCreate) || !ep("cuGreenCtxStreamCreate", pStream) || !ep("cuDeviceGet", pDevGet)) return 1;
  CUdevice dev; DR(pDevGet(&dev, 0));
  CUdevResource all; DR(pGetRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("device SMs: %u\n", all.sm.smCount);
  for (unsigned want : {8u, 16u}) {
    CUdevResource grp[1], rem; unsigned nb = 1;
    CUresult r = pSplit(grp, &nb, &all, &rem, 0, want);
    printf("split min %u: rc %d groups %u group SMs %u remainder %u\n", want, (int)r, nb, grp[0].sm.smCount, rem.sm.smCount);
  }
  CUdevResource grp[1], rem; unsigned nb = 1;
  DR(pSplit(grp, &nb, &all, &rem, 0, 8));
  CUdevResourceDesc dA, dB; DR(pDesc(&dA, &grp[0], 1)); DR(pDesc(&dB, &rem, 1));
  CUgreenCtx gA, gB; DR(pCreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM)); DR(pCreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  int lo = 0, hi = 0; RT(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CUstream sA, sB, sB2; DR(pStream(&sA, gA, CU_STREAM_NON_BLOCKING, hi)); DR(pStream(&sB, gB, CU_STREAM_NON_BLOCKING, 0)); DR(pStream(&sB2, gB, CU_STREAM_NON_BLOCKING, lo));
  cudaStream_t s0; RT(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
  int *oA, *oB, *oC; RT(cudaMalloc(&oA, 4096 * 4)); RT(cudaMalloc(&oB, 4096 * 4)); RT(cudaMalloc(&oC, 64));
  std::vector<int> hA(64), hB(1024), hC(8);
  // 1. runtime launches into the green streams
  who<<<64, 128, 0, (cudaStream_t)sA>>>(oA, 1000); RT(cudaGetLastError());
  who<<<1024, 128, 0, (cudaStream_t)sB>>>(oB, 1000); RT(cudaGetLastError());
  RT(cudaDeviceSynchronize());
  RT(cudaMemcpy(hA.data(), oA, 64 * 4, cudaMemcpyDeviceToHost)); RT(cudaMemcpy(hB.data(), oB, 1024 * 4, cudaMemcpyDeviceToHost));
  show("partition A (runtime launch)", hA); show("partition B (runtime launch)", hB);
  // 2. cluster of 8 with 152 KB dynamic smem in partition A
  RT(cudaFuncSetAttribute(who_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 152 * 1024));
  who_cluster<<<8, 256, 152 * 1024, (cudaStream_t)sA>>>(oC);
  cudaError_t ce = cudaGetLastError(); printf("cluster launch in A: %s\n", cudaGetErrorString(ce));
  ce = cudaDeviceSynchronize(); printf("cluster sync: %s\n", cudaGetErrorString(ce));
  if (ce == cudaSuccess) { RT(cudaMemcpy(hC.data(), oC, 32, cudaMemcpyDeviceToHost)); show("cluster in A", hC); }
  // 3. latency of a small A kernel while B is saturated by a long low-priority kernel
  cudaEvent_t e0, e1, e2; RT(cudaEventCreate(&e0)); RT(cudaEventCreate(&e1)); RT(cudaEventCreate(&e2));
  for (int mode = 0; mode < 2; ++mode) {
    cudaStream_t big = mode == 0 ? (cudaStream_t)sB2 : s0;   // 0: saturate partition B, 1: saturate the whole device (primary ctx)
    RT(cudaEventRecord(e0, (cudaStream_t)sA));
    who<<<4096, 256, 80 * 1024 * 0, big>>>(oB, 200000);   // ~100 us per CTA
    who<<<1, 128, 0, (cudaStream_t)sA>>>(oA, 1000);
    RT(cudaEventRecord(e1, (cudaStream_t)sA));
    RT(cudaEventRecord(e2, big));
    RT(cudaDeviceSynchronize());
    float ta = 0, tb = 0; RT(cudaEventElapsedTime(&ta, e0, e1)); RT(cudaEventElapsedTime(&tb, e0, e2));
    printf("mode %d: small kernel in A done after %.1f us, big kernel after %.1f us\n", mode, ta * 1e3, tb * 1e3);
  }
  // 4. stream capture on a primary-context stream, fork into the green streams, join, replay
  cudaEvent_t f, jA, jB; RT(cudaEventCreateWithFlags(&f, cudaEventDisableTiming)); RT(cudaEventCreateWithFlags(&jA, cudaEventDisableTiming)); RT(cudaEventCreateWithFlags(&jB, cudaEventDisableTiming));
  cudaGraph_t graph; cudaGraphExec_t exec;
  RT(cudaMemset(oA, 0xff, 64 * 4)); RT(cudaMemset(oB, 0xff, 1024 * 4));
  ce = cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal); printf("begin capture: %s\n", cudaGetErrorString(ce));
  RT(cudaEventRecord(f, s0));
  ce = cudaStreamWaitEvent((cudaStream_t)sA, f, 0); printf("green A waits fork (capture): %s\n", cudaGetErrorString(ce));
  ce = cudaStreamWaitEvent((cudaStream_t)sB, f, 0); printf("green B waits fork (capture): %s\n", cudaGetErrorString(ce));
  who<<<64, 128, 0, (cudaStream_t)sA>>>(oA, 1000); printf("capture launch A: %s\n", cudaGetErrorString(cudaGetLastError()));
  who<<<1024, 128, 0, (cudaStream_t)sB>>>(oB, 1000); printf("capture launch B: %s\n", cudaGetErrorString(cudaGetLastError()));
  cudaEventRecord(jA, (cudaStream_t)sA); cudaEventRecord(jB, (cudaStream_t)sB);
  cudaStreamWaitEvent(s0, jA, 0); cudaStreamWaitEvent(s0, jB, 0);
  ce = cudaStreamEndCapture(s0, &graph); printf("end capture: %s\n", cudaGetErrorString(ce));
  if (ce == cudaSuccess) {
    ce = cudaGraphInstantiate(&exec, graph, 0); printf("instantiate: %s\n", cudaGetErrorString(ce));
    if (ce == cudaSuccess) {
      ce = cudaGraphLaunch(exec, s0); printf("graph launch: %s\n", cudaGetErrorString(ce));
      ce = cudaDeviceSynchronize(); printf("graph sync: %s\n", cudaGetErrorString(ce));
      RT(cudaMemcpy(hA.data(), oA, 64 * 4, cudaMemcpyDeviceToHost)); RT(cudaMemcpy(hB.data(), oB, 1024 * 4, cudaMemcpyDeviceToHost));
      show("graph replay: partition A", hA); show("graph replay: partition B", hB);
    }
  }
  printf("done\n");
  return 0;
}
