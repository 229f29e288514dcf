// Micro-benchmark: non-tensor FP32 / FP64 FMA throughput of the device (scalar FFMA vs packed FFMA2 =
// fma.rn.f32x2, sm_100+), to know how close the Dslash kernels are to being issue-bound.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_ffma(float* out, float a, float b, int iters) {
  float x[ILP];
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_ffma2(float* out, float a, float b, int iters) {
  unsigned long long x[ILP];
  unsigned long long av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
  for (int i = 0; i < ILP; i++) {
    float v = threadIdx.x * 1e-3f + i;
    asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v));
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(av), "l"(bv));
  }
  float s = 0;
  for (int i = 0; i < ILP; i++) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dfma(double* out, double a, double b, int iters) {
  double x[ILP];
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  const int threads = 256, blocks = sms * 8, iters = 4096, ILP = 8;
  float* o;
  cudaMalloc(&o, sizeof(double) * threads * blocks);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    for (int r = 0; r < 10; r++) k_ffma<ILP><<<blocks, threads>>>(o, 1.0001f, 1e-6f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = 10.0 * 2.0 * ILP * iters * threads * blocks / (ms * 1e-3) / 1e12;
    if (rep) printf("FFMA  scalar : %.2f TFLOP/s\n", tf);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; r++) k_ffma2<ILP><<<blocks, threads>>>(o, 1.0001f, 1e-6f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    tf = 10.0 * 4.0 * ILP * iters * threads * blocks / (ms * 1e-3) / 1e12;
    if (rep) printf("FFMA2 packed : %.2f TFLOP/s\n", tf);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; r++) k_dfma<ILP><<<blocks, threads>>>((double*)o, 1.0001, 1e-6, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    tf = 10.0 * 2.0 * ILP * iters * threads * blocks / (ms * 1e-3) / 1e12;
    if (rep) printf("DFMA  scalar : %.2f TFLOP/s\n", tf);
  }
  printf("SMs %d clock %d kHz\n", sms, p.clockRate);
  return 0;
}
