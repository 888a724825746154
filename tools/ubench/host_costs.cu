// Host-side cost of the runtime calls behind the likelihood graph replay: pinned allocations, capture,
// instantiate, launch, destroy.   nvcc -O2 -arch=sm_100a -o host_costs host_costs.cu && ./host_costs
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* p) { if (threadIdx.x == 0) p[0] += 1.0; }
static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  double* d; cudaMalloc(&d, 1 << 20);
  for (int rep = 0; rep < 3; ++rep) {
    double t0 = now();
    double* h[3];
    for (int i = 0; i < 3; ++i) cudaMallocHost(&h[i], 64);
    double t1 = now();
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    cudaMemcpyAsync(d, h[0], 24, cudaMemcpyHostToDevice, st);
    for (int i = 0; i < 4; ++i) k<<<1, 32, 0, st>>>(d);
    cudaMemsetAsync(d + 64, 0, 4096, st);
    cudaMemcpyAsync(h[1], d, 24, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(h[2], d, 12, cudaMemcpyDeviceToHost, st);
    cudaStreamEndCapture(st, &g);
    double t2 = now();
    cudaGraphInstantiate(&ge, g, 0);
    double t3 = now();
    cudaGraphDestroy(g);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    double t4 = now();
    for (int i = 0; i < 100; ++i) { cudaGraphLaunch(ge, st); cudaStreamSynchronize(st); }
    double t5 = now();
    cudaGraphExecDestroy(ge);
    double t6 = now();
    for (int i = 0; i < 3; ++i) cudaFreeHost(h[i]);
    double t7 = now();
    double* dd; cudaMalloc(&dd, 1 << 22); double t8 = now(); cudaFree(dd); double t9 = now();
    printf("3x cudaMallocHost %.0f us | capture %.0f us | instantiate %.0f us | first launch+sync %.0f us | launch+sync %.1f us each | "
           "exec destroy %.0f us | 3x cudaFreeHost %.0f us | cudaMalloc 4MB %.0f us | cudaFree %.0f us\n",
           t1 - t0, t2 - t1, t3 - t2, t4 - t3, (t5 - t4) / 100, t6 - t5, t7 - t6, t8 - t7, t9 - t8);
  }
  return 0;
}
