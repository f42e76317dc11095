// deepfluids_b200 -- small fp32 GEMM for the fully connected layers outside the generator's K <= 16 case.
//
// slim.fully_connected (ops.py:23-24) appears on this path as: the generator FC (K = 3 / 16: dfl_fc_fwd), the encoder FC
// on the channel-blocked concat tensor (dfl_enc_fc_*), and -- at the ops level / in the latent-space MLP of arch=nn
// (model.py:218-224: linear 2*filters, linear filters, linear onum) -- plain [B,K] x [K,N] products of any size.  Those are
// tiny (<= 1024 x 1024 x batch) and far from any roofline that matters for the step; this is a plain 64x64x16 shared-memory
// tiled SIMT kernel, fp32 FMA in the reference's arithmetic type, with split-K (atomics) when M*N alone cannot fill the chip
// (not in deterministic mode, dfl_set_deterministic: one CTA per output tile walks all of K in order).
//   C[M,N] = op(A)[M,K] * op(B)[K,N] (+ bias[N]) (+ C if accumulate)
//   A: row-major [M,K], or [K,M] when transA;  B: row-major [K,N], or [N,K] when transB.
#include "dfl_common.cuh"

namespace dfl {

constexpr int GT = 64, GK = 16;

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B, long long sbk,
                long long sbn, const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K, int kchunk,
                int accumulate, int atomic) {
  __shared__ float sA[GK][GT + 1], sB[GK][GT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  const int k_lo = blockIdx.z * kchunk, k_hi = min(K, k_lo + kchunk);
  float acc[4][4] = {};
  for (int k0 = k_lo; k0 < k_hi; k0 += GK) {
    for (int i = threadIdx.x; i < GK * GT; i += 256) {
      const int kk = i / GT, mm = i % GT;      // (coalesced when the operand's unit stride runs along m / n)
      const int k = k0 + kk;
      sA[kk][mm] = (k < k_hi && m0 + mm < M) ? __ldg(A + (m0 + mm) * sam + k * sak) : 0.f;
      sB[kk][mm] = (k < k_hi && n0 + mm < N) ? __ldg(B + k * sbk + (n0 + mm) * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m >= M || n >= N) continue;
      float v = acc[i][j];
      if (bias && blockIdx.z == 0) v += __ldg(bias + n);
      float* c = C + static_cast<size_t>(m) * N + n;
      if (atomic) atomicAdd(c, v);
      else *c = accumulate ? *c + v : v;
    }
}

int gemm_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int transA, int transB,
             int accumulate, cudaStream_t st) {
  DFL_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "gemm_f32: null operand or empty shape (%d,%d,%d)", M, N, K);
  const long long sam = transA ? 1 : K, sak = transA ? M : 1, sbk = transB ? 1 : N, sbn = transB ? K : 1;
  const int gx = (N + GT - 1) / GT, gy = (M + GT - 1) / GT;
  int nsplit = 1;
  size_t det_bytes = 0;
  const bool det = deterministic_workspace(&det_bytes) != nullptr;     // dfl_set_deterministic: no atomics -> no split-K
  if (!det && gx * gy < num_sms() && K >= 4096) nsplit = std::min((K + 2047) / 2048, std::max(1, 2 * num_sms() / (gx * gy)));
  int kchunk = ((K + nsplit - 1) / nsplit + GK - 1) / GK * GK;
  nsplit = (K + kchunk - 1) / kchunk;
  if (nsplit > 1 && !accumulate) DFL_CUDA_OK(cudaMemsetAsync(C, 0, static_cast<size_t>(M) * N * sizeof(float), st));
  gemm_f32_kernel<<<dim3(gx, gy, nsplit), 256, 0, st>>>(A, sam, sak, B, sbk, sbn, bias, C, M, N, K, kchunk, accumulate,
                                                       nsplit > 1 ? 1 : 0);
  DFL_LAUNCH_OK("gemm_f32_kernel");
  return DFL_OK;
}

// out[n] = sum_m x[m][n]   (BiasAddGrad of a fully connected layer); one thread per column, coalesced over n
__global__ void colsum_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += __ldg(x + static_cast<size_t>(m) * N + n);
  out[n] = s;
}

int colsum_f32(const float* x, float* out, int M, int N, cudaStream_t st) {
  DFL_REQUIRE(x && out && M > 0 && N > 0, "colsum_f32: null operand or empty shape");
  colsum_f32_kernel<<<(N + 127) / 128, 128, 0, st>>>(x, out, M, N);
  DFL_LAUNCH_OK("colsum_f32_kernel");
  return DFL_OK;
}

}  // namespace dfl
