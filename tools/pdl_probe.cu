// pdl_probe.cu -- what a chain of dependent kernels costs on this GPU (build: nvcc -arch=sm_100a -O3 -o pdl_probe pdl_probe.cu)
//
// A step of the batch-1 GEMV stack is 224 kernels, each of which needs the output of the one before: the hand-off
// between two kernels of a stream bounds the whole thing from below.  This probe measures, inside a captured CUDA
// graph of CHAIN kernels of `grid` CTAs:
//   period     wall time per kernel of the chain (CUDA events around graph replays)
//   entry->wait   time a CTA spends between its first instruction and the return of griddepcontrol.wait
//   exit->wait    time between the LAST CTA of kernel i leaving its work loop and kernel i+1's wait returning
// for plain stream order and for programmatic dependent launch, with a small and a large shared-memory footprint
// (large = the successor cannot be co-resident and only starts when a slot frees up).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}

// ts[k*4 + {0: first CTA entry, 1: first wait return, 2: last work end}] via atomicMin / atomicMax
__global__ void k_link(unsigned long long *ts, int k, int work_ns, float *sink, int pdl) {
  extern __shared__ unsigned char smem[];
  const unsigned long long t0 = gtime();
  if (pdl) asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x == 0) atomicMin(&ts[k * 4 + 0], t0);
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const unsigned long long t1 = gtime();
  if (threadIdx.x == 0) atomicMin(&ts[k * 4 + 1], t1);
  float acc = 0.f;
  while (gtime() - t1 < (unsigned long long)work_ns) acc += 1.0f;
  if (threadIdx.x == 0) {
    smem[0] = (unsigned char)acc;
    atomicMax(&ts[k * 4 + 2], gtime());
    if (acc < 0.f) sink[0] = acc + smem[0];
  }
}

static void run(const char *name, int grid, int threads, int smem, int work_ns, int pdl, int chain) {
  unsigned long long *ts;
  float *sink;
  cudaMalloc(&ts, chain * 4 * sizeof(unsigned long long));
  cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(k_link, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaStream_t s;
  cudaStreamCreate(&s);
  auto init = [&]() {
    std::vector<unsigned long long> h(chain * 4);
    for (int i = 0; i < chain; i++) { h[i * 4] = h[i * 4 + 1] = ~0ull; h[i * 4 + 2] = 0; h[i * 4 + 3] = 0; }
    cudaMemcpy(ts, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  };
  init();
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int k = 0; k < chain; k++) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = smem; lc.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr; lc.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&lc, k_link, ts, k, work_ns, sink, pdl);
  }
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  for (int i = 0; i < 3; i++) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 20;
  cudaEventRecord(e0, s);
  for (int i = 0; i < reps; i++) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  init();
  cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  std::vector<unsigned long long> h(chain * 4);
  cudaMemcpy(h.data(), ts, h.size() * 8, cudaMemcpyDeviceToHost);
  double ew = 0, xw = 0, ee = 0;
  int n = 0;
  for (int k = 8; k < chain - 1; k++) {
    ew += (double)(h[k * 4 + 1] - h[k * 4 + 0]);
    xw += (double)((long long)h[(k + 1) * 4 + 1] - (long long)h[k * 4 + 2]);
    ee += (double)((long long)h[(k + 1) * 4 + 0] - (long long)h[k * 4 + 2]);
    n++;
  }
  cudaError_t err = cudaGetLastError();
  printf("{\"probe\": \"%s\", \"grid\": %d, \"threads\": %d, \"smem\": %d, \"work_ns\": %d, \"pdl\": %d, \"period_us\": %.3f, "
         "\"entry_to_wait_us\": %.3f, \"lastexit_to_nextwait_us\": %.3f, \"lastexit_to_nextentry_us\": %.3f, \"err\": \"%s\"}\n",
         name, grid, threads, smem, work_ns, pdl, ms * 1e3 / (reps * chain), ew / n * 1e-3, xw / n * 1e-3, ee / n * 1e-3,
         cudaGetErrorString(err));
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaFree(ts); cudaFree(sink); cudaStreamDestroy(s);
}

int main() {
  const int chain = 200;
  for (int work : {0, 2000}) {
    run("plain_small", 148, 256, 1024, work, 0, chain);
    run("pdl_small", 148, 256, 1024, work, 1, chain);
    run("pdl_small_2perSM", 296, 256, 1024, work, 1, chain);
    run("pdl_half_sm", 148, 512, 100 * 1024, work, 1, chain);        // successor co-resident (2 x 100 KB fit)
    run("pdl_full_sm", 148, 512, 200 * 1024, work, 1, chain);        // successor waits for the slot
    run("pdl_2x_half", 296, 256, 100 * 1024, work, 1, chain);        // two CTAs per SM fill it: successor waits for a slot
    run("plain_full_sm", 148, 512, 200 * 1024, work, 0, chain);
  }
  return 0;
}
