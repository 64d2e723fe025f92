// Kernel dispatch throughput of a CUDA graph with B parallel branches of tiny dependent kernels (what a chunk farm of
// small matrices looks like to the launch hardware): microseconds per kernel node.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__global__ void tiny(int* p) { if (threadIdx.x == 0 && p) atomicAdd(p, 1); }
__global__ void spin(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
int main() {
  for (int B : {1, 8, 32, 64}) for (int mode = 0; mode < 2; ++mode) {
    const int L = 400;
    std::vector<cudaStream_t> st(B + 1);
    std::vector<cudaEvent_t> ev(B + 1);
    for (auto& s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st[B], cudaStreamCaptureModeThreadLocal);
    cudaEventRecord(ev[B], st[B]);
    for (int b = 0; b < B; ++b) {
      cudaStreamWaitEvent(st[b], ev[B], 0);
      for (int i = 0; i < L; ++i) { if (mode == 0) tiny<<<1, 32, 0, st[b]>>>(nullptr); else spin<<<4, 256, 0, st[b]>>>(4000); }
      cudaEventRecord(ev[b], st[b]);
      cudaStreamWaitEvent(st[B], ev[b], 0);
    }
    cudaStreamEndCapture(st[B], &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaGraphLaunch(ge, st[B]); cudaStreamSynchronize(st[B]);
    cudaEventRecord(e0, st[B]); cudaGraphLaunch(ge, st[B]); cudaEventRecord(e1, st[B]); cudaStreamSynchronize(st[B]);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%2d branches x %d %s kernels: %.2f ms -> %.3f us per kernel node overall, %.2f us per node along a branch\n", B, L,
           mode ? "2-us (4 CTAs)" : "empty", ms, ms * 1e3 / (B * L), ms * 1e3 / L);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  }
  return 0;
}
