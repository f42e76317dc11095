// Which named-barrier patterns does `compute-sanitizer --tool synccheck` flag?  Every kernel below is a CORRECT use of
// bar.sync / bar.arrive with an explicit thread count (PTX ISA: "bar.sync a, b": b threads, a multiple of the warp size,
// whole warps participate).  Build + run:  nvcc -arch=sm_100a -o /tmp/probe tools/synccheck_probe.cu &&
//   for k in A B C D E F G; do compute-sanitizer --tool synccheck /tmp/probe $k; done
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void bsync(int id, int n) { __syncwarp(); asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void barrive(int id, int n) { __syncwarp(); asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// A: warps 0,1 meet at barrier 1 from the SAME call site; warps 2,3 idle until the final __syncthreads
__global__ void kA(int* o) {
  const int w = threadIdx.x >> 5;
  if (w < 2) bsync(1, 64);
  __syncthreads();
  if (threadIdx.x == 0) o[0] = 1;
}
// B: warps 0 and 1 meet at barrier 1 from DIFFERENT call sites (role-split code)
__global__ void kB(int* o) {
  const int w = threadIdx.x >> 5;
  if (w == 0) { o[1] = 1; bsync(1, 64); }
  else if (w == 1) { o[2] = 2; o[3] = 3; bsync(1, 64); }
  __syncthreads();
}
// C: producer bar.arrive / consumer bar.sync
__global__ void kC(int* o) {
  const int w = threadIdx.x >> 5;
  if (w == 0) { o[4] = 1; barrive(2, 64); }
  else if (w == 1) { bsync(2, 64); o[5] = o[4]; }
  __syncthreads();
}
// D: as B, repeated in loops with role-dependent trip structure (the shape of the fused kernel's store rounds)
__global__ void kD(int* o) {
  const int w = threadIdx.x >> 5;
  if (w == 0) { for (int i = 0; i < 4; ++i) { o[6] = i; bsync(3, 64); } }
  else if (w == 1) { for (int t = 0; t < 2; ++t) for (int h = 0; h < 2; ++h) { o[7] = h; bsync(3, 64); } }
  __syncthreads();
}
// E: as B while the other warps have already EXITED
__global__ void kE(int* o) {
  const int w = threadIdx.x >> 5;
  if (w >= 2) return;
  if (w == 0) { o[8] = 1; bsync(1, 64); } else { o[9] = 2; bsync(1, 64); }
}
// F: as B, but both roles reach the barrier through ONE non-inlined function: a single bar.sync instruction in the binary
__device__ __noinline__ void bsync_shared(int id, int n) { __syncwarp(); asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__global__ void kF(int* o) {
  const int w = threadIdx.x >> 5;
  if (w == 0) { o[10] = 1; bsync_shared(1, 64); }
  else if (w == 1) { o[11] = 2; o[12] = 3; bsync_shared(1, 64); }
  __syncthreads();
}
// G: arrive / sync, each through its own non-inlined function
__device__ __noinline__ void barrive_shared(int id, int n) { __syncwarp(); asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__global__ void kG(int* o) {
  const int w = threadIdx.x >> 5;
  if (w == 0) { o[13] = 1; barrive_shared(2, 64); }
  else if (w == 1) { bsync_shared(2, 64); o[14] = o[13]; }
  __syncthreads();
}
int main(int argc, char** argv) {
  int* o; cudaMalloc(&o, 64 * sizeof(int));
  const char* names[] = {"A same-site", "B split-site", "C arrive/sync", "D loops", "E others-exited", "F split-site via one noinline function",
                         "G arrive/sync via noinline functions"};
  void (*ks[])(int*) = {kA, kB, kC, kD, kE, kF, kG};
  const int i = argc > 1 ? argv[1][0] - 'A' : 0;      // one kernel per process: a synccheck error is sticky for the context
  printf("== kernel %s\n", names[i]); fflush(stdout);
  ks[i]<<<1, 128>>>(o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("   -> %s\n", cudaGetErrorString(e)); fflush(stdout);
  return 0;
}
