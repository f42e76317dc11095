// deepfluids_b200 -- HBM-bound edge layers and element-wise glue of the generator train step (SIMT kernels).
//
//   fc_fwd / fc_bwd            slim.fully_connected, activation None      (reference ops.py:23-24, model.py:19,61)
//   lastconv_{fwd,dgrad,wgrad} the 128 -> 1/2/3 channel output conv       (model.py:42,84): reads/writes 128-ch
//                              activations once per voxel -> memory/FMA-bound, not a tensor-core shape (N <= 3)
//   pool_mask                  adjoint of nearest x2 upsample (sum over the 2x2(x2) children, ops.py:75-91)
//                              fused with the leaky-ReLU derivative of the layer below (ops.py:9-10)
//   bias_grad                  column sums of dL/d(pre-activation)
//   pack_conv_weights          fp32 TF-layout master weights -> bf16 K-major GEMM operands (fwd + dgrad)
//   adam_step                  tf.train.AdamOptimizer update (trainer.py:160-162), one fused pass over a flat buffer
#include "dfl_common.cuh"

namespace dfl {

// =============================================================================================
// FC:  out[b,n] = sum_k z[b,k] W[k,n] + bias[n]      z fp32 [B,K], W fp32 [K,N] (TF [in,out]), out bf16 [B,N]
// =============================================================================================
constexpr int FC_MAXK = 16, FC_MAXB = 64;

template <typename TO>
__global__ void fc_fwd_kernel(const float* __restrict__ z, const float* __restrict__ W, const float* __restrict__ bias,
                              TO* __restrict__ out, int B, int K, int N) {
  __shared__ float sz[FC_MAXB * FC_MAXK];
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) sz[i] = z[i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float w[FC_MAXK];
#pragma unroll
  for (int k = 0; k < FC_MAXK; ++k) w[k] = (k < K) ? W[static_cast<size_t>(k) * N + n] : 0.f;
  const float bn = bias[n];
  for (int b = 0; b < B; ++b) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < FC_MAXK; ++k)
      if (k < K) a = fmaf(sz[b * K + k], w[k], a);
    stf(out + static_cast<size_t>(b) * N + n, a + bn);
  }
}

// dW[k,n] = sum_b z[b,k] dOut[b,n];  db[n] = sum_b dOut[b,n]   (written, not accumulated)
template <typename TI>
__global__ void fc_bwd_kernel(const float* __restrict__ z, const TI* __restrict__ dout, float* __restrict__ dW,
                              float* __restrict__ db, int B, int K, int N) {
  __shared__ float sz[FC_MAXB * FC_MAXK];
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) sz[i] = z[i];
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc[FC_MAXK];
#pragma unroll
  for (int k = 0; k < FC_MAXK; ++k) acc[k] = 0.f;
  float sb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float g = ldf(dout + static_cast<size_t>(b) * N + n);
    sb += g;
#pragma unroll
    for (int k = 0; k < FC_MAXK; ++k)
      if (k < K) acc[k] = fmaf(sz[b * K + k], g, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < FC_MAXK; ++k)
    if (k < K) dW[static_cast<size_t>(k) * N + n] = acc[k];
  db[n] = sb;
}

int fc_fwd(const float* z, const float* W, const float* bias, void* out, int B, int K, int N, int out_dtype,
           cudaStream_t st) {
  DFL_REQUIRE(K <= FC_MAXK && B <= FC_MAXB, "fc_fwd: K <= %d and B <= %d required (got K=%d B=%d)", FC_MAXK, FC_MAXB,
              K, B);
  const int threads = 256, grid = (N + threads - 1) / threads;
  if (out_dtype == DT_BF16)
    fc_fwd_kernel<<<grid, threads, 0, st>>>(z, W, bias, static_cast<__nv_bfloat16*>(out), B, K, N);
  else
    fc_fwd_kernel<<<grid, threads, 0, st>>>(z, W, bias, static_cast<float*>(out), B, K, N);
  DFL_LAUNCH_OK("fc_fwd_kernel");
  return DFL_OK;
}

int fc_bwd(const float* z, const void* dout, float* dW, float* db, int B, int K, int N, int dout_dtype,
           cudaStream_t st) {
  DFL_REQUIRE(K <= FC_MAXK && B <= FC_MAXB, "fc_bwd: K <= %d and B <= %d required (got K=%d B=%d)", FC_MAXK, FC_MAXB,
              K, B);
  const int threads = 256, grid = (N + threads - 1) / threads;
  if (dout_dtype == DT_BF16)
    fc_bwd_kernel<<<grid, threads, 0, st>>>(z, static_cast<const __nv_bfloat16*>(dout), dW, db, B, K, N);
  else
    fc_bwd_kernel<<<grid, threads, 0, st>>>(z, static_cast<const float*>(dout), dW, db, B, K, N);
  DFL_LAUNCH_OK("fc_bwd_kernel");
  return DFL_OK;
}

// =============================================================================================
// last conv 128 -> COUT (1..3), k = 3, SAME, no activation.  x bf16 [B,D,H,W,128]; W fp32 TF layout
// [taps][128][COUT]; out fp32 [B,D,H,W,COUT].   One warp per run of 8 consecutive x voxels; lane l owns input
// channels 4l..4l+3; the x-1..x+8 neighbour slices are loaded once per (dz,dy) and reused by the three dx taps.
// =============================================================================================
constexpr int LC_RUN = 8;

struct LcParams {
  int B, D, H, W, kd;   // kd = 3 (3D) or 1 (2D)
  int runs_per_row, nruns;
};

__device__ __forceinline__ void ld_bf16x4(const __nv_bfloat16* p, float (&f)[4]) {
  const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
  f[0] = __uint_as_float(q.x << 16);
  f[1] = __uint_as_float(q.x & 0xFFFF0000u);
  f[2] = __uint_as_float(q.y << 16);
  f[3] = __uint_as_float(q.y & 0xFFFF0000u);
}

template <int COUT>
__global__ void __launch_bounds__(128)
lastconv_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                    float* __restrict__ out, LcParams p) {
  extern __shared__ float sw[];   // [taps][COUT][128]
  const int ntaps = p.kd * 9;
  for (int i = threadIdx.x; i < ntaps * 128 * COUT; i += blockDim.x) {
    const int co = i % COUT, ci = (i / COUT) % 128, t = i / (COUT * 128);
    sw[(t * COUT + co) * 128 + ci] = W[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int run = wglobal; run < p.nruns; run += nw) {
    int r = run;
    const int xr = (r % p.runs_per_row) * LC_RUN; r /= p.runs_per_row;
    const int y = r % p.H; r /= p.H;
    const int z = r % p.D;
    const int b = r / p.D;
    float acc[LC_RUN][COUT];
#pragma unroll
    for (int j = 0; j < LC_RUN; ++j)
#pragma unroll
      for (int c = 0; c < COUT; ++c) acc[j][c] = 0.f;
    for (int dz = 0; dz < p.kd; ++dz) {
      const int zz = z + dz - (p.kd >> 1);
      if (zz < 0 || zz >= p.D) continue;
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
        if (yy < 0 || yy >= p.H) continue;
        const __nv_bfloat16* row = x + ((static_cast<size_t>(b) * p.D + zz) * p.H + yy) * p.W * 128 + lane * 4;
        float xin[LC_RUN + 2][4];
#pragma unroll
        for (int j = 0; j < LC_RUN + 2; ++j) {
          const int xx = xr + j - 1;
          if (xx >= 0 && xx < p.W) ld_bf16x4(row + static_cast<size_t>(xx) * 128, xin[j]);
          else xin[j][0] = xin[j][1] = xin[j][2] = xin[j][3] = 0.f;
        }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int t = (dz * 3 + dy) * 3 + dx;
          float w[COUT][4];
#pragma unroll
          for (int c = 0; c < COUT; ++c) {
            const float4 q = *reinterpret_cast<const float4*>(&sw[(t * COUT + c) * 128 + lane * 4]);
            w[c][0] = q.x; w[c][1] = q.y; w[c][2] = q.z; w[c][3] = q.w;
          }
#pragma unroll
          for (int j = 0; j < LC_RUN; ++j)
#pragma unroll
            for (int c = 0; c < COUT; ++c)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[j][c] = fmaf(xin[j + dx][e], w[c][e], acc[j][c]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < LC_RUN; ++j)
#pragma unroll
      for (int c = 0; c < COUT; ++c) acc[j][c] = warp_sum(acc[j][c]);
    if (lane < LC_RUN * COUT) {
      const int j = lane / COUT, c = lane % COUT;
      float v = 0.f;
#pragma unroll
      for (int jj = 0; jj < LC_RUN; ++jj)
#pragma unroll
        for (int cc = 0; cc < COUT; ++cc)
          if (jj == j && cc == c) v = acc[jj][cc];
      const int xx = xr + j;
      if (xx < p.W)
        out[(((static_cast<size_t>(b) * p.D + z) * p.H + y) * p.W + xx) * COUT + c] = v + bias[c];
    }
  }
}

// dX[p, ci] = sum_{t,co} dOut[p - (t-1), co] W[t, ci, co];  optional second output dX * lrelu'(mask_src)
template <int COUT>
__global__ void __launch_bounds__(128)
lastconv_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ W,
                      const __nv_bfloat16* __restrict__ mask_src, __nv_bfloat16* __restrict__ dx,
                      __nv_bfloat16* __restrict__ dx_masked, LcParams p) {
  extern __shared__ float sw[];   // [taps][COUT][128]
  const int ntaps = p.kd * 9;
  for (int i = threadIdx.x; i < ntaps * 128 * COUT; i += blockDim.x) {
    const int co = i % COUT, ci = (i / COUT) % 128, t = i / (COUT * 128);
    sw[(t * COUT + co) * 128 + ci] = W[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int run = wglobal; run < p.nruns; run += nw) {
    int r = run;
    const int xr = (r % p.runs_per_row) * LC_RUN; r /= p.runs_per_row;
    const int y = r % p.H; r /= p.H;
    const int z = r % p.D;
    const int b = r / p.D;
    float acc[LC_RUN][4];
#pragma unroll
    for (int j = 0; j < LC_RUN; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    for (int dz = 0; dz < p.kd; ++dz) {
      const int zz = z - (dz - (p.kd >> 1));
      if (zz < 0 || zz >= p.D) continue;
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = y - (dy - 1);
        if (yy < 0 || yy >= p.H) continue;
        const float* row = dout + ((static_cast<size_t>(b) * p.D + zz) * p.H + yy) * p.W * COUT;
        // gradient values at x positions xr-1 .. xr+8 (same for all lanes: broadcast loads)
        float g[LC_RUN + 2][COUT];
#pragma unroll
        for (int j = 0; j < LC_RUN + 2; ++j) {
          const int xx = xr + j - 1;
#pragma unroll
          for (int c = 0; c < COUT; ++c) g[j][c] = (xx >= 0 && xx < p.W) ? __ldg(row + static_cast<size_t>(xx) * COUT + c) : 0.f;
        }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int t = (dz * 3 + dy) * 3 + dx;
          float w[COUT][4];
#pragma unroll
          for (int c = 0; c < COUT; ++c) {
            const float4 q = *reinterpret_cast<const float4*>(&sw[(t * COUT + c) * 128 + lane * 4]);
            w[c][0] = q.x; w[c][1] = q.y; w[c][2] = q.z; w[c][3] = q.w;
          }
          // output voxel xr+j receives dOut[x - (dx-1)] = g[j + 1 - (dx-1)] = g[j + 2 - dx]
#pragma unroll
          for (int j = 0; j < LC_RUN; ++j)
#pragma unroll
            for (int c = 0; c < COUT; ++c)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[j][e] = fmaf(g[j + 2 - dx][c], w[c][e], acc[j][e]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < LC_RUN; ++j) {
      const int xx = xr + j;
      if (xx >= p.W) continue;
      const size_t off = ((((static_cast<size_t>(b) * p.D + z) * p.H + y) * p.W + xx) * 128) + lane * 4;
      if (dx) {
        uint2 o;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[j][0], acc[j][1]), h1 = __floats2bfloat162_rn(acc[j][2], acc[j][3]);
        o.x = *reinterpret_cast<uint32_t*>(&h0);
        o.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(dx + off) = o;
      }
      if (dx_masked) {
        float m[4];
        ld_bf16x4(mask_src + off, m);
        uint2 o;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[j][0] * lrelu_grad_from_out(m[0]), acc[j][1] * lrelu_grad_from_out(m[1]));
        __nv_bfloat162 h1 = __floats2bfloat162_rn(acc[j][2] * lrelu_grad_from_out(m[2]), acc[j][3] * lrelu_grad_from_out(m[3]));
        o.x = *reinterpret_cast<uint32_t*>(&h0);
        o.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(dx_masked + off) = o;
      }
    }
  }
}

// dW[t][ci][co] += sum_p x[p+t-1][ci] dOut[p][co];  db[co] += sum_p dOut[p][co].
// Block = 4 voxel-lanes x kd warps; warp (s, dz) walks input voxels q of sub-slab s and scatters x[q] into the
// 9 (dy,dx) taps of plane dz:  dW[t] += x[q] * dOut[q - (t-1)].  Lane l owns ci = 4l..4l+3.
template <int COUT>
__global__ void __launch_bounds__(384)
lastconv_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ dout, float* __restrict__ dW,
                      float* __restrict__ db, LcParams p) {
  __shared__ float red[3][9][COUT][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int dzi = warp % p.kd, sub = warp / p.kd;
  const int nsub = (blockDim.x >> 5) / p.kd;
  for (int i = threadIdx.x; i < 3 * 9 * COUT * 128; i += blockDim.x) (&red[0][0][0][0])[i] = 0.f;
  __syncthreads();
  float acc[9][COUT][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[t][c][0] = acc[t][c][1] = acc[t][c][2] = acc[t][c][3] = 0.f;
  float bsum[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) bsum[c] = 0.f;
  const size_t nvox = static_cast<size_t>(p.B) * p.D * p.H * p.W;
  const int dz = dzi - (p.kd >> 1) + ((p.kd == 1) ? 0 : 0);
  for (size_t q = static_cast<size_t>(blockIdx.x) * nsub + sub; q < nvox; q += static_cast<size_t>(gridDim.x) * nsub) {
    const int xq = q % p.W, yq = (q / p.W) % p.H, zq = (q / (static_cast<size_t>(p.W) * p.H)) % p.D;
    float xv[4];
    ld_bf16x4(x + q * 128 + lane * 4, xv);
    // tap (dz,dy,dx) pairs x[q] with dOut at p = q - (tap - 1)
    const int zp = zq - dz;
    if (zp >= 0 && zp < p.D) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yp = yq - (dy - 1);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int xp = xq - (dx - 1);
          if (yp >= 0 && yp < p.H && xp >= 0 && xp < p.W) {
            const float* g = dout + (q + (static_cast<ptrdiff_t>(zp - zq) * p.H + (yp - yq)) * p.W + (xp - xq)) * COUT;
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
              const float gv = __ldg(g + c);
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[dy * 3 + dx][c][e] = fmaf(xv[e], gv, acc[dy * 3 + dx][c][e]);
            }
          }
        }
      }
    }
    if (dzi == 0 && lane == 0) {
#pragma unroll
      for (int c = 0; c < COUT; ++c) bsum[c] += __ldg(dout + q * COUT + c);
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < COUT; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(&red[dzi][t][c][lane * 4 + e], acc[t][c][e]);
  __syncthreads();
  const int ntaps = p.kd * 9;
  for (int i = threadIdx.x; i < ntaps * COUT * 128; i += blockDim.x) {
    const int ci = i % 128, c = (i / 128) % COUT, t = i / (128 * COUT);
    atomicAdd(&dW[(static_cast<size_t>(t) * 128 + ci) * COUT + c], red[t / 9][t % 9][c][ci]);
  }
  if (dzi == 0 && lane == 0) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) atomicAdd(&db[c], bsum[c]);
  }
}

static void lc_plan(const int64_t* dims, int nd, LcParams& p) {
  p.B = static_cast<int>(dims[0]);
  p.D = nd == 3 ? static_cast<int>(dims[1]) : 1;
  p.H = static_cast<int>(dims[nd - 1]);
  p.W = static_cast<int>(dims[nd]);
  p.kd = nd == 3 ? 3 : 1;
  p.runs_per_row = (p.W + LC_RUN - 1) / LC_RUN;
  p.nruns = p.B * p.D * p.H * p.runs_per_row;
}

template <int COUT>
static int lastconv_dispatch(int op, const void* a, const void* b, const void* c, void* o0, void* o1,
                             const LcParams& p, cudaStream_t st) {
  const int ntaps = p.kd * 9;
  const size_t smem = static_cast<size_t>(ntaps) * COUT * 128 * sizeof(float);
  const int grid = std::min((p.nruns + 3) / 4, num_sms() * 8);
  if (op == 0) {
    static bool set = false;
    if (!set && smem > 48 * 1024) {
      DFL_CUDA_OK(cudaFuncSetAttribute(lastconv_fwd_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      set = true;
    }
    lastconv_fwd_kernel<COUT><<<grid, 128, smem, st>>>(static_cast<const __nv_bfloat16*>(a), static_cast<const float*>(b),
                                                       static_cast<const float*>(c), static_cast<float*>(o0), p);
  } else if (op == 1) {
    static bool set = false;
    if (!set && smem > 48 * 1024) {
      DFL_CUDA_OK(cudaFuncSetAttribute(lastconv_dgrad_kernel<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      set = true;
    }
    lastconv_dgrad_kernel<COUT><<<grid, 128, smem, st>>>(static_cast<const float*>(a), static_cast<const float*>(b),
                                                         static_cast<const __nv_bfloat16*>(c),
                                                         static_cast<__nv_bfloat16*>(o0),
                                                         static_cast<__nv_bfloat16*>(o1), p);
  } else {
    const int wgrid = num_sms() * 2;
    lastconv_wgrad_kernel<COUT><<<wgrid, 4 * p.kd * 32, 0, st>>>(static_cast<const __nv_bfloat16*>(a),
                                                                 static_cast<const float*>(b), static_cast<float*>(o0),
                                                                 static_cast<float*>(o1), p);
  }
  DFL_LAUNCH_OK("lastconv_kernel");
  return DFL_OK;
}

// op 0: fwd(a=x, b=W, c=bias -> o0=out f32)   op 1: dgrad(a=dout, b=W, c=mask_src|null -> o0=dx|null, o1=dx_masked|null)
// op 2: wgrad(a=x, b=dout -> o0=dW (+=), o1=db (+=))
int lastconv(int op, const void* a, const void* b, const void* c, void* o0, void* o1, const int64_t* dims, int nd,
             int cout, cudaStream_t st) {
  DFL_REQUIRE(nd == 2 || nd == 3, "lastconv: ndim must be 2 or 3");
  DFL_REQUIRE(cout >= 1 && cout <= 3, "lastconv: Cout must be 1..3 (got %d)", cout);
  LcParams p{};
  lc_plan(dims, nd, p);
  switch (cout) {
    case 1: return lastconv_dispatch<1>(op, a, b, c, o0, o1, p, st);
    case 2: return lastconv_dispatch<2>(op, a, b, c, o0, o1, p, st);
    default: return lastconv_dispatch<3>(op, a, b, c, o0, o1, p, st);
  }
}

// =============================================================================================
// pool_mask: ds[b,z,y,x,c] = sum over the 2x2(x2) children of g (fine grid);  dmasked = ds * lrelu'(mask_src)
// =============================================================================================
__global__ void pool_mask_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ mask_src,
                                 __nv_bfloat16* __restrict__ ds, __nv_bfloat16* __restrict__ dmasked, int B, int D,
                                 int H, int W, int zr) {
  // coarse dims D,H,W; fine dims D*zr, 2H, 2W; 16 threads per voxel (8 channels each)
  const size_t n = static_cast<size_t>(B) * D * H * W * 16;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int q = idx & 15;
    size_t v = idx >> 4;
    const int x = v % W; v /= W;
    const int y = v % H; v /= H;
    const int z = v % D;
    const int b = v / D;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int D2 = D * zr, H2 = 2 * H, W2 = 2 * W;
    for (int a = 0; a < zr; ++a)
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const size_t pos2 = ((static_cast<size_t>(b) * D2 + (z * zr + a)) * H2 + (2 * y + e)) * W2 + (2 * x + f);
          const uint4 qv = __ldg(reinterpret_cast<const uint4*>(g + pos2 * 128 + q * 8));
          const uint32_t w[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            s[2 * k] += __uint_as_float(w[k] << 16);
            s[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
          }
        }
    const size_t off = (idx >> 4) * 128 + q * 8;
    auto pack = [](float a, float b2) {
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b2);
      return *reinterpret_cast<uint32_t*>(&h);
    };
    if (ds) *reinterpret_cast<uint4*>(ds + off) = make_uint4(pack(s[0], s[1]), pack(s[2], s[3]), pack(s[4], s[5]), pack(s[6], s[7]));
    if (dmasked) {
      const uint4 mv = __ldg(reinterpret_cast<const uint4*>(mask_src + off));
      const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
      float m[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        m[2 * k] = s[2 * k] * lrelu_grad_from_out(__uint_as_float(w[k] << 16));
        m[2 * k + 1] = s[2 * k + 1] * lrelu_grad_from_out(__uint_as_float(w[k] & 0xFFFF0000u));
      }
      *reinterpret_cast<uint4*>(dmasked + off) = make_uint4(pack(m[0], m[1]), pack(m[2], m[3]), pack(m[4], m[5]), pack(m[6], m[7]));
    }
  }
}

int pool_mask(const void* g, const void* mask_src, void* ds, void* dmasked, const int64_t* cdims, int nd,
              cudaStream_t st) {
  const int B = cdims[0], D = nd == 3 ? cdims[1] : 1, H = cdims[nd - 1], W = cdims[nd];
  const size_t n = static_cast<size_t>(B) * D * H * W * 16;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  pool_mask_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(mask_src),
                                         static_cast<__nv_bfloat16*>(ds), static_cast<__nv_bfloat16*>(dmasked), B, D, H, W,
                                         nd == 3 ? 2 : 1);
  DFL_LAUNCH_OK("pool_mask_kernel");
  return DFL_OK;
}

// =============================================================================================
// bias_grad: db[c] += sum_pos d[pos][c]      d bf16 [npos][128]
// =============================================================================================
__global__ void __launch_bounds__(256) bias_grad_kernel(const __nv_bfloat16* __restrict__ d, float* __restrict__ db,
                                                        size_t npos) {
  __shared__ float red[16][128];
  const int q = threadIdx.x & 15, r = threadIdx.x >> 4;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t pos = static_cast<size_t>(blockIdx.x) * 16 + r; pos < npos; pos += static_cast<size_t>(gridDim.x) * 16) {
    const uint4 qv = __ldg(reinterpret_cast<const uint4*>(d + pos * 128 + q * 8));
    const uint32_t w[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s[2 * k] += __uint_as_float(w[k] << 16);
      s[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[r][q * 8 + k] = s[k];
  __syncthreads();
  if (threadIdx.x < 128) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) t += red[k][threadIdx.x];
    atomicAdd(db + threadIdx.x, t);
  }
}

int bias_grad(const void* d, float* db, size_t npos, cudaStream_t st) {
  const int grid = static_cast<int>(std::min<size_t>((npos + 15) / 16, static_cast<size_t>(num_sms()) * 8));
  bias_grad_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(d), db, npos);
  DFL_LAUNCH_OK("bias_grad_kernel");
  return DFL_OK;
}

// =============================================================================================
// pack_conv_weights: W fp32 [taps][Cin][Cout] (TF DHWIO/HWIO)  ->
//    wf bf16 [Cout][taps*Cin]          (fwd B operand: row n = co, K = tap*Cin + ci)
//    wd bf16 [Cin][taps*Cout]          (dgrad B operand: row n = ci, K = tap'*Cout + co, tap' = flipped tap)
// =============================================================================================
__global__ void pack_conv_weights_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ wf,
                                         __nv_bfloat16* __restrict__ wd, int taps, int cin, int cout) {
  const int n = taps * cin * cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int co = i % cout, ci = (i / cout) % cin, t = i / (cout * cin);
    const __nv_bfloat16 v = __float2bfloat16_rn(W[i]);
    if (wf) wf[static_cast<size_t>(co) * taps * cin + t * cin + ci] = v;
    if (wd) wd[static_cast<size_t>(ci) * taps * cout + (taps - 1 - t) * cout + co] = v;
  }
}

int pack_conv_weights(const float* W, void* wf, void* wd, int taps, int cin, int cout, cudaStream_t st) {
  const int n = taps * cin * cout;
  pack_conv_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(W, static_cast<__nv_bfloat16*>(wf),
                                                            static_cast<__nv_bfloat16*>(wd), taps, cin, cout);
  DFL_LAUNCH_OK("pack_conv_weights_kernel");
  return DFL_OK;
}

// =============================================================================================
// adam_step (TF semantics): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t * m / (sqrt(v) + eps)
//   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) is computed by the host and passed in.  grad_scale folds 1/world.
// =============================================================================================
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, size_t n, float lr_t, const float* __restrict__ lr_t_dev,
                                 float b1, float b2, float eps, float grad_scale) {
  if (lr_t_dev) lr_t = *lr_t_dev;      // CUDA-graph friendly: the step size lives in device memory
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}
__global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, size_t n, float lr,
                                const float* __restrict__ lr_dev, float grad_scale) {
  if (lr_dev) lr = *lr_dev;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    p[i] -= lr * g[i] * grad_scale;
}

int adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr_t, const float* lr_t_dev, float b1,
              float b2, float eps, float grad_scale, cudaStream_t st) {
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  if (m && v)
    adam_step_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n, lr_t, lr_t_dev, b1, b2, eps, grad_scale);
  else
    sgd_step_kernel<<<grid, 256, 0, st>>>(p, g, n, lr_t, lr_t_dev, grad_scale);
  DFL_LAUNCH_OK("adam_step_kernel");
  return DFL_OK;
}

// fp32 -> bf16 cast (targets, inputs)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ a, __nv_bfloat16* __restrict__ o, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    o[i] = __float2bfloat16_rn(a[i]);
}
int cast_f32_bf16(const float* a, void* o, size_t n, cudaStream_t st) {
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  cast_f32_bf16_kernel<<<grid, 256, 0, st>>>(a, static_cast<__nv_bfloat16*>(o), n);
  DFL_LAUNCH_OK("cast_kernel");
  return DFL_OK;
}

}  // namespace dfl
