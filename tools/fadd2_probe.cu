// Does sm_100a's packed FADD2 (add.rn.f32x2) double fp32 add throughput?  8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_scalar(float* out, int n) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  const float b = out[0];
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = a[i] + b;
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x + 1] = s;
}
__global__ void k_packed(float* out, int n) {
  unsigned long long a[8];
  for (int i = 0; i < 8; ++i) {
    float2 v = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 1e-3f + i + 8);
    a[i] = *reinterpret_cast<unsigned long long*>(&v);
  }
  float2 bb = make_float2(out[0], out[0]);
  const unsigned long long b = *reinterpret_cast<unsigned long long*>(&bb);
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&a[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x + 1] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, (148 * 4 * 512 + 1) * sizeof(float));
  cudaMemset(out, 0, 4);
  const int n = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    float ms;
    cudaEventRecord(e0); k_scalar<<<148 * 4, 512>>>(out, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double adds = 148.0 * 4 * 512 * 16 * n;
    printf("scalar FADD : %.3f ms  %.2f T adds/s\n", ms, adds / ms / 1e9);
    cudaEventRecord(e0); k_packed<<<148 * 4, 512>>>(out, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("packed FADD2: %.3f ms  %.2f T adds/s\n", ms, adds / ms / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
