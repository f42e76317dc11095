// umma_probe -- which shared-memory element does tcgen05.mma read for (row m, k) when the smem descriptor's start
// address is NOT 1024-byte aligned and/or SBO is not a multiple of 1024 (128B swizzle)?
// The answer decides whether convolution taps can be expressed as shifted descriptor windows over ONE resident
// activation halo brick (DESIGN.md "tap windows").  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../deep-fluids_b200/csrc/dfl_common.cuh"

using namespace dfl;

struct Cfg {
  int mn_major;      // 0: A K-major (rows = M), 1: A MN-major (rows = K)
  int start_row;     // descriptor start = base + start_row*128 (+ kbyte)
  int kbyte;         // extra byte offset inside the row (K advance for K-major): 0 or 32/64/96
  int sbo, lbo;      // bytes
  int base_off;      // descriptor base_offset field
};

__global__ void probe_kernel(Cfg c, int pass, float* out /*[128][16]*/) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int ROWS = 400;
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);                 // ROWS x 128 B, TMA-style SW128 image
  __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(smem + ROWS * 128);   // 16 x 128 B (1024-aligned: 400*128=51200)
  for (int i = threadIdx.x; i < ROWS * 64; i += blockDim.x) {
    const int r = i / 64, col = i % 64;
    const int chunk = col / 8, e = col % 8;
    const int pc = chunk ^ (r & 7);
    const float v = (pass == 0) ? static_cast<float>(r % 256) : static_cast<float>(col);
    A[r * 64 + pc * 8 + e] = __float2bfloat16(v);
  }
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
    const int n = i / 64, col = i % 64;
    const int chunk = col / 8, e = col % 8;
    const int pc = chunk ^ (n & 7);
    Bm[n * 64 + pc * 8 + e] = __float2bfloat16((col == n) ? 1.f : 0.f);
  }
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tptr, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 16, c.mn_major, 0);
    const uint64_t da = umma_desc_sw128(smem_u32(smem) + c.start_row * 128 + c.kbyte, c.lbo, c.sbo, c.base_off);
    const uint64_t db = umma_desc_sw128(smem_u32(Bm), 16, 1024, 0);
    umma_bf16(tm, da, db, idesc, 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rr[32];
  tmem_ld_32x32(tm + (static_cast<uint32_t>(warp * 32) << 16), rr);
  tmem_ld_wait();
  for (int k = 0; k < 16; ++k) out[(warp * 32 + lane) * 16 + k] = __uint_as_float(rr[k]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 32);
}

int main() {
  std::vector<Cfg> cfgs = {
      {0, 0, 0, 1024, 16, 0},     // canonical
      {0, 0, 32, 1024, 16, 0},    // K advance (known good)
      {0, 1, 0, 1024, 16, 0},     // start shifted by one row, base_offset 0
      {0, 1, 0, 1024, 16, 1},     // ... base_offset 1
      {0, 3, 0, 1024, 16, 0},
      {0, 3, 0, 1024, 16, 3},
      {0, 0, 0, 1280, 16, 0},     // SBO = 10 rows
      {0, 3, 0, 1280, 16, 0},
      {0, 3, 0, 1280, 16, 3},
      {0, 11, 64, 2304, 16, 0},   // 18-row pitch, shifted start, K advance
      {0, 11, 64, 2304, 16, 3},
      {1, 0, 0, 1024, 16384, 0},  // MN-major canonical (two 64-ch blocks 16 KB apart; only block 0 + garbage read)
      {1, 1, 0, 1024, 16384, 0},  // MN-major, K window shifted by one row
      {1, 1, 0, 1024, 16384, 1},
      {1, 5, 0, 1280, 16384, 0},  // second 8-row K group 10 rows later
      {1, 5, 0, 1280, 16384, 5},
  };
  float* d;
  cudaMalloc(&d, 128 * 16 * 4);
  std::vector<float> h0(128 * 16), h1(128 * 16);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (size_t ci = 0; ci < cfgs.size(); ++ci) {
    const Cfg& c = cfgs[ci];
    for (int pass = 0; pass < 2; ++pass) {
      probe_kernel<<<1, 128, 100 * 1024>>>(c, pass, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("cfg %zu: CUDA error %s\n", ci, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(pass == 0 ? h0.data() : h1.data(), d, 128 * 16 * 4, cudaMemcpyDeviceToHost);
    }
    // model ABS: swizzle is a function of the absolute smem address => logical (row, col) = affine in (m, k)
    int ok = 0, tot = 0;
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 16; ++k) {
        int er, ec;
        if (!c.mn_major) {
          er = c.start_row + (m % 8) + (m / 8) * (c.sbo / 128);
          ec = c.kbyte / 2 + k;
        } else {
          if (m >= 64) continue;    // second 64-channel block lies outside the image
          er = c.start_row + (k % 8) + (k / 8) * (c.sbo / 128);
          ec = m;
        }
        ++tot;
        if (static_cast<int>(h0[m * 16 + k]) == er % 256 && static_cast<int>(h1[m * 16 + k]) == ec) ++ok;
      }
    printf("cfg %2zu major=%s start_row=%2d kbyte=%2d sbo=%4d base_off=%d : ABS-model match %4d/%4d |", ci,
           c.mn_major ? "MN" : "K ", c.start_row, c.kbyte, c.sbo, c.base_off, ok, tot);
    const int ms[5] = {0, 1, 7, 8, 9};
    for (int mi = 0; mi < 5; ++mi) {
      const int m = ms[mi];
      printf(" m%d:(r%d,c%d)(r%d,c%d)", m, (int)h0[m * 16 + 0], (int)h1[m * 16 + 0], (int)h0[m * 16 + 9],
             (int)h1[m * 16 + 9]);
    }
    printf("\n");
  }
  return 0;
}
