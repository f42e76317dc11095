// deepfluids_b200 -- batch-norm + activation and dropout for the latent-space MLP of arch=nn
// (reference model.py:218-224: linear(2*filters) -> batch_norm(act=elu) -> dropout, linear(filters) -> batch_norm(act=elu) ->
// dropout, linear(onum); ops.py:26-36 batch_norm = slim.batch_norm(decay=momentum, epsilon, scale=True, fused=True,
// updates_collections=None, is_training=train, activation_fn=act)).  Tensors are [M = batch, N = features] fp32: a few
// hundred KB -- these kernels are about semantics, not rooflines.
//   training:  mean_n = mean_m x, var_n = mean_m (x - mean)^2 (biased, used to normalise),
//              y = act(gamma * (x - mean) * rsqrt(var + eps) + beta),
//              moving_mean = decay * moving_mean + (1 - decay) * mean,
//              moving_var  = decay * moving_var  + (1 - decay) * var * M / (M - 1)     (fused batch norm hands slim the
//                                                                                      Bessel-corrected variance)
//   inference: y = act(gamma * (x - moving_mean) * rsqrt(moving_var + eps) + beta)
//   act: 0 none, 1 leaky-ReLU(0.2) (ops.batch_norm's default), 2 ELU (tf.nn.elu, NN's default)
#include "dfl_common.cuh"

namespace dfl {

__device__ __forceinline__ float mlp_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.2f * v);
  if (act == 2) return v > 0.f ? v : expm1f(v);
  return v;
}
// derivative of the activation expressed on its OUTPUT y (sign(y) == sign(pre); ELU': 1 for pre > 0, exp(pre) = y + 1 else)
__device__ __forceinline__ float mlp_act_grad(float y, int act) {
  if (act == 1) return y >= 0.f ? 1.f : 0.2f;
  if (act == 2) return y > 0.f ? 1.f : y + 1.f;
  return 1.f;
}

constexpr int BN_TX = 32, BN_TY = 8;

__global__ void __launch_bounds__(BN_TX* BN_TY)
bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float* __restrict__ mmean, float* __restrict__ mvar, float* __restrict__ y, float* __restrict__ save_mean,
                  float* __restrict__ save_rstd, int M, int N, float eps, float decay, int training, int act) {
  __shared__ float red[BN_TY][BN_TX];
  const int n = blockIdx.x * BN_TX + threadIdx.x;
  const bool ok = n < N;
  float mean, rstd;
  if (training) {
    float s = 0.f;
    if (ok) for (int m = threadIdx.y; m < M; m += BN_TY) s += x[static_cast<size_t>(m) * N + n];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    s = 0.f;
    for (int k = 0; k < BN_TY; ++k) s += red[k][threadIdx.x];
    mean = s / M;
    __syncthreads();
    float q = 0.f;
    if (ok) for (int m = threadIdx.y; m < M; m += BN_TY) { const float d = x[static_cast<size_t>(m) * N + n] - mean; q += d * d; }
    red[threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    q = 0.f;
    for (int k = 0; k < BN_TY; ++k) q += red[k][threadIdx.x];
    const float var = q / M;
    rstd = rsqrtf(var + eps);
    if (ok && threadIdx.y == 0) {
      save_mean[n] = mean;
      save_rstd[n] = rstd;
      mmean[n] = decay * mmean[n] + (1.f - decay) * mean;
      mvar[n] = decay * mvar[n] + (1.f - decay) * (M > 1 ? var * M / (M - 1) : var);
    }
  } else {
    mean = ok ? mmean[n] : 0.f;
    rstd = ok ? rsqrtf(mvar[n] + eps) : 0.f;
  }
  if (!ok) return;
  const float g = gamma[n] * rstd, b = beta[n] - mean * g;
  for (int m = threadIdx.y; m < M; m += BN_TY) {
    const size_t i = static_cast<size_t>(m) * N + n;
    y[i] = mlp_act(fmaf(x[i], g, b), act);
  }
}

// dz = dy * act'(y);  dbeta = sum dz;  dgamma = sum dz * xhat;  dx = gamma * rstd / M * (M dz - dbeta - xhat dgamma)
__global__ void __launch_bounds__(BN_TX* BN_TY)
bn_act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                  const float* __restrict__ gamma, const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                  float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int N, int act) {
  __shared__ float r0[BN_TY][BN_TX], r1[BN_TY][BN_TX];
  const int n = blockIdx.x * BN_TX + threadIdx.x;
  const bool ok = n < N;
  const float mean = ok ? save_mean[n] : 0.f, rstd = ok ? save_rstd[n] : 0.f;
  float sb = 0.f, sg = 0.f;
  if (ok)
    for (int m = threadIdx.y; m < M; m += BN_TY) {
      const size_t i = static_cast<size_t>(m) * N + n;
      const float dz = dy[i] * mlp_act_grad(y[i], act);
      sb += dz;
      sg += dz * (x[i] - mean) * rstd;
    }
  r0[threadIdx.y][threadIdx.x] = sb;
  r1[threadIdx.y][threadIdx.x] = sg;
  __syncthreads();
  sb = sg = 0.f;
  for (int k = 0; k < BN_TY; ++k) { sb += r0[k][threadIdx.x]; sg += r1[k][threadIdx.x]; }
  if (!ok) return;
  if (threadIdx.y == 0) { dbeta[n] = sb; dgamma[n] = sg; }
  if (!dx) return;
  const float c = gamma[n] * rstd / M;
  for (int m = threadIdx.y; m < M; m += BN_TY) {
    const size_t i = static_cast<size_t>(m) * N + n;
    const float dz = dy[i] * mlp_act_grad(y[i], act);
    dx[i] = c * (M * dz - sb - (x[i] - mean) * rstd * sg);
  }
}

int bn_act_fwd(const float* x, const float* gamma, const float* beta, float* mmean, float* mvar, float* y, float* save_mean,
               float* save_rstd, int M, int N, float eps, float decay, int training, int act, cudaStream_t st) {
  DFL_REQUIRE(x && gamma && beta && mmean && mvar && y && M > 0 && N > 0, "bn_act_fwd: null tensor or empty shape");
  DFL_REQUIRE(!training || (save_mean && save_rstd), "bn_act_fwd: training mode needs save_mean / save_rstd");
  DFL_REQUIRE(act >= 0 && act <= 2, "bn_act_fwd: act must be 0 (none), 1 (lrelu) or 2 (elu)");
  bn_act_fwd_kernel<<<(N + BN_TX - 1) / BN_TX, dim3(BN_TX, BN_TY), 0, st>>>(x, gamma, beta, mmean, mvar, y, save_mean, save_rstd,
                                                                          M, N, eps, decay, training, act);
  DFL_LAUNCH_OK("bn_act_fwd_kernel");
  return DFL_OK;
}

int bn_act_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* save_mean,
               const float* save_rstd, float* dx, float* dgamma, float* dbeta, int M, int N, int act, cudaStream_t st) {
  DFL_REQUIRE(x && y && dy && gamma && save_mean && save_rstd && dgamma && dbeta && M > 0 && N > 0, "bn_act_bwd: null tensor");
  bn_act_bwd_kernel<<<(N + BN_TX - 1) / BN_TX, dim3(BN_TX, BN_TY), 0, st>>>(x, y, dy, gamma, save_mean, save_rstd, dx, dgamma,
                                                                          dbeta, M, N, act);
  DFL_LAUNCH_OK("bn_act_bwd_kernel");
  return DFL_OK;
}

// slim.dropout(x, keep_prob, is_training) (model.py:220-221; NOTE the reference passes its `dropout=0.1` argument as
// KEEP probability): y = x * mask / keep_prob, mask ~ Bernoulli(keep_prob) from a counter-based generator keyed by
// (seed, offset + element index), so the backward pass re-creates the mask by calling the same entry point on dy.
__device__ __forceinline__ float dropout_uniform(unsigned long long seed, unsigned long long ctr) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (ctr + 1);      // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<float>(z >> 40) * (1.0f / 16777216.0f);            // 24 random bits -> [0, 1)
}
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, float keep, unsigned long long seed,
                               unsigned long long offset) {
  const float inv = 1.f / keep;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    y[i] = dropout_uniform(seed, offset + i) < keep ? x[i] * inv : 0.f;
}
int dropout(const float* x, float* y, size_t n, float keep, unsigned long long seed, unsigned long long offset, cudaStream_t st) {
  DFL_REQUIRE(x && y && keep > 0.f && keep <= 1.f, "dropout: null tensor or keep_prob outside (0, 1]");
  if (n == 0) return DFL_OK;
  dropout_kernel<<<static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 8)), 256, 0, st>>>(x, y, n, keep, seed, offset);
  DFL_LAUNCH_OK("dropout_kernel");
  return DFL_OK;
}

}  // namespace dfl
