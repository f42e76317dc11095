// deepfluids_b200 -- HBM-bound edge layers and element-wise glue of the generator train step (SIMT kernels).
//
//   fc_fwd / fc_bwd            slim.fully_connected, activation None      (reference ops.py:23-24, model.py:19,61)
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
// pool_mask: ds[b,z,y,x,c] = sum over the 2x2(x2) children of g (fine grid);  dmasked = ds * lrelu'(mask_src)
// =============================================================================================
__global__ void pool_mask_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ addend,
                                 const __nv_bfloat16* __restrict__ mask_src, __nv_bfloat16* __restrict__ ds,
                                 __nv_bfloat16* __restrict__ dmasked, int B, int D, int H, int W, int zr) {
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
    if (addend) {        // a coarse-grid term of the same gradient (phase-decomposed upsample-conv: its data gradient)
      const uint4 av = __ldg(reinterpret_cast<const uint4*>(addend + off));
      const uint32_t w[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s[2 * k] += __uint_as_float(w[k] << 16);
        s[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
      }
    }
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
              cudaStream_t st, const void* addend) {
  const int B = cdims[0], D = nd == 3 ? cdims[1] : 1, H = cdims[nd - 1], W = cdims[nd];
  const size_t n = static_cast<size_t>(B) * D * H * W * 16;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  pool_mask_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(addend),
                                         static_cast<const __nv_bfloat16*>(mask_src),
                                         static_cast<__nv_bfloat16*>(ds), static_cast<__nv_bfloat16*>(dmasked), B, D, H, W,
                                         nd == 3 ? 2 : 1);
  DFL_LAUNCH_OK("pool_mask_kernel");
  return DFL_OK;
}

// =============================================================================================
// pack_phase_weights: operands of the PHASE-DECOMPOSED upsample-conv.
// conv3(upsample_x2(s)) (model.py:76-79 followed by :67-69) evaluated at the fine voxel 2p + r (r = phase, one bit per
// axis) only ever sees TWO distinct coarse voxels per axis -- r = 0: s[p-1] (tap 0) and s[p] (taps 1 + 2); r = 1: s[p]
// (taps 0 + 1) and s[p+1] (tap 2) -- so it equals 2^nd convolutions with 2^nd taps on the coarse tensor whose weights are
// sums of the layer's 3^nd taps: 8/27 of the dense FLOPs in 3D, 4/9 in 2D; the zero padding carries over unchanged.
//   w   fp32 TF layout [3^nd][cin][cout]
//   wf  bf16 [P][cout][T*cin]      forward operand of phase r (rows co, columns (tap o, ci))          P = T = 2^nd
//   wd  bf16 [cin][P*T*cout]       data-gradient operand (rows ci, columns (phase r, tap o, co))
// phase / tap indices: most significant bit = first spatial axis (z in 3D); tap bit o_a = 0 -> the lower coarse offset.
// =============================================================================================
__global__ void pack_phase_weights_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ wf,
                                          __nv_bfloat16* __restrict__ wd, int nd, int cin, int cout) {
  const int P = 1 << nd, T = 1 << nd;
  const size_t n = static_cast<size_t>(P) * T * cin * cout;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = idx % cout;
    size_t v = idx / cout;
    const int ci = v % cin; v /= cin;
    const int o = v % T;
    const int r = static_cast<int>(v / T);
    // per axis: the fine taps [lo, hi] that fold onto coarse offset bit o_a in phase bit r_a
    float acc = 0.f;
    int lo[3], hi[3];
    for (int a = 0; a < nd; ++a) {
      const int ra = (r >> (nd - 1 - a)) & 1, oa = (o >> (nd - 1 - a)) & 1;
      if (ra == 0) { lo[a] = oa ? 1 : 0; hi[a] = oa ? 2 : 0; }
      else { lo[a] = oa ? 2 : 0; hi[a] = oa ? 2 : 1; }
    }
    if (nd == 2) { lo[2] = hi[2] = 0; }
    for (int t0 = lo[0]; t0 <= hi[0]; ++t0)
      for (int t1 = lo[1]; t1 <= hi[1]; ++t1)
        for (int t2 = lo[2]; t2 <= hi[2]; ++t2) {
          const int t = nd == 3 ? (t0 * 3 + t1) * 3 + t2 : t0 * 3 + t1;
          acc += __ldg(W + (static_cast<size_t>(t) * cin + ci) * cout + co);
        }
    const __nv_bfloat16 h = __float2bfloat16_rn(acc);
    wf[(static_cast<size_t>(r) * cout + co) * (static_cast<size_t>(T) * cin) + static_cast<size_t>(o) * cin + ci] = h;
    wd[static_cast<size_t>(ci) * (static_cast<size_t>(P) * T * cout) + (static_cast<size_t>(r) * T + o) * cout + co] = h;
  }
}

int pack_phase_weights(const float* W, void* wf, void* wd, int nd, int cin, int cout, cudaStream_t st) {
  DFL_REQUIRE(nd == 2 || nd == 3, "pack_phase_weights: ndim must be 2 or 3");
  DFL_REQUIRE(W && wf && wd, "pack_phase_weights: null tensor");
  const size_t n = static_cast<size_t>(1 << nd) * (1 << nd) * cin * cout;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  pack_phase_weights_kernel<<<grid, 256, 0, st>>>(W, static_cast<__nv_bfloat16*>(wf), static_cast<__nv_bfloat16*>(wd), nd, cin, cout);
  DFL_LAUNCH_OK("pack_phase_weights_kernel");
  return DFL_OK;
}

// =============================================================================================
// gather_stride2: coarse[b,z,y,x,:] = fine[b,2z,2y,2x,:]  (bf16, 128 channels).  The up-sampled tensor x0 = upscale(s) read at
// every second voxel IS s: the coarse operand of the phase-decomposed weight gradient, exactly as the forward pass read it.
// =============================================================================================
__global__ void gather_stride2_kernel(const __nv_bfloat16* __restrict__ fine, __nv_bfloat16* __restrict__ coarse, int B,
                                      int D, int H, int W, int zr) {
  const size_t n = static_cast<size_t>(B) * D * H * W * 16;      // 16 threads per coarse voxel (8 channels = 16 bytes each)
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int q = idx & 15;
    size_t v = idx >> 4;
    const int x = v % W; v /= W;
    const int y = v % H; v /= H;
    const int z = v % D;
    const int b = v / D;
    const size_t pos2 = ((static_cast<size_t>(b) * (D * zr) + z * zr) * (2 * H) + 2 * y) * (2 * W) + 2 * x;
    reinterpret_cast<uint4*>(coarse)[idx] = __ldg(reinterpret_cast<const uint4*>(fine + pos2 * 128) + q);
  }
}

int gather_stride2(const void* fine, void* coarse, const int64_t* cdims, int nd, cudaStream_t st) {
  const int B = cdims[0], D = nd == 3 ? cdims[1] : 1, H = cdims[nd - 1], W = cdims[nd];
  const size_t n = static_cast<size_t>(B) * D * H * W * 16;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  gather_stride2_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(fine), static_cast<__nv_bfloat16*>(coarse), B, D, H,
                                             W, nd == 3 ? 2 : 1);
  DFL_LAUNCH_OK("gather_stride2_kernel");
  return DFL_OK;
}

// =============================================================================================
// phase_wgrad_fold: weight gradient of the phase-decomposed upsample-conv, folded back onto the layer's 3^nd taps.
// dfl_phase_wgrad (the tensor-core weight-gradient kernel run as a 4^nd-tap stride-2 correlation between dY on the fine grid
// and the layer's COARSE input s) leaves T[k][co][ci] = sum_q dY[2q + k - 1][co] * s[q][ci], k in {0..3}^nd, i.e. the gradients
// of the pre-summed phase weights, transposed.  Per axis the phase (r, off) that reads fine offset k - 1 = r - 2 off carries
// the taps {0} (k=3), {1,2} (k=1), {0,1} (k=2), {2} (k=0); hence tap t receives k in {3,2} (t=0), {1,2} (t=1), {1,0} (t=2):
//     dW[t][ci][co] += sum_{k_a in M[t_a]} T[k][co][ci].
// =============================================================================================
__global__ void phase_wgrad_fold_kernel(const float* __restrict__ T, float* __restrict__ dw, int nd, int cin, int cout) {
  const int ntap = nd == 3 ? 27 : 9;
  const size_t n = static_cast<size_t>(ntap) * cin * cout;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = idx % cin;                 // ci fastest: coalesced reads of T[k][co][ci]
    size_t v = idx / cin;
    const int co = v % cout;
    const int t = static_cast<int>(v / cout);
    int ta[3] = {0, 0, 0};
    if (nd == 3) { ta[0] = t / 9; ta[1] = (t / 3) % 3; ta[2] = t % 3; } else { ta[0] = t / 3; ta[1] = t % 3; }
    float acc = 0.f;
    const int nk = 1 << nd;
    for (int m = 0; m < nk; ++m) {
      int k = 0;
      for (int a = 0; a < nd; ++a) {
        const int sel = (m >> (nd - 1 - a)) & 1;
        const int ka = ta[a] == 0 ? (sel ? 2 : 3) : (ta[a] == 1 ? (sel ? 2 : 1) : (sel ? 0 : 1));
        k = k * 4 + ka;
      }
      acc += __ldg(T + (static_cast<size_t>(k) * cout + co) * cin + ci);
    }
    dw[(static_cast<size_t>(t) * cin + ci) * cout + co] += acc;
  }
}

int phase_wgrad_fold(const float* T, float* dw, int nd, int cin, int cout, cudaStream_t st) {
  DFL_REQUIRE(nd == 2 || nd == 3, "phase_wgrad_fold: ndim must be 2 or 3");
  const size_t n = static_cast<size_t>(nd == 3 ? 27 : 9) * cin * cout;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  phase_wgrad_fold_kernel<<<grid, 256, 0, st>>>(T, dw, nd, cin, cout);
  DFL_LAUNCH_OK("phase_wgrad_fold_kernel");
  return DFL_OK;
}

// =============================================================================================
// bias_grad: db[c] += sum_pos d[pos][c]      d bf16 [npos][128]
// =============================================================================================
// `partial` != nullptr (deterministic mode): block sums go to partial[blockIdx.x][128]; bias_grad_reduce_kernel adds them
// to db in block order
__global__ void __launch_bounds__(256) bias_grad_kernel(const __nv_bfloat16* __restrict__ d, float* __restrict__ db,
                                                        size_t npos, float* __restrict__ partial) {
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
    if (partial) partial[static_cast<size_t>(blockIdx.x) * 128 + threadIdx.x] = t;
    else atomicAdd(db + threadIdx.x, t);
  }
}
__global__ void __launch_bounds__(128) bias_grad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ db, int n) {
  float t = 0.f;
  for (int i = 0; i < n; ++i) t += partial[static_cast<size_t>(i) * 128 + threadIdx.x];
  db[threadIdx.x] += t;
}

int bias_grad(const void* d, float* db, size_t npos, cudaStream_t st) {
  const int grid = static_cast<int>(std::min<size_t>((npos + 15) / 16, static_cast<size_t>(num_sms()) * 8));
  size_t wb = 0;
  float* ws = deterministic_workspace(&wb);
  DFL_REQUIRE(!ws || static_cast<size_t>(grid) * 128 * sizeof(float) <= wb, "bias_grad: deterministic workspace too small");
  bias_grad_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(d), db, npos, ws);
  DFL_LAUNCH_OK("bias_grad_kernel");
  if (ws) {
    bias_grad_reduce_kernel<<<1, 128, 0, st>>>(ws, db, grid);
    DFL_LAUNCH_OK("bias_grad_reduce_kernel");
  }
  return DFL_OK;
}

// =============================================================================================
// pack_conv_weights: W fp32 [taps][Cin][Cout] (TF DHWIO/HWIO)  ->
//    wf bf16 [Cout][taps*Cin]          (fwd B operand: row n = co, K = tap*Cin + ci)
//    wd bf16 [Cin][taps*Cout]          (dgrad B operand: row n = ci, K = tap'*Cout + co, tap' = flipped tap)
// =============================================================================================
// `cin_ld` >= cin is the padded input-channel count of the forward operand (rows of zeros stay untouched), used by the
// encoder's first conv whose 2/3 input channels are zero-padded to 128.
__global__ void pack_conv_weights_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ wf,
                                         __nv_bfloat16* __restrict__ wd, int taps, int cin, int cout, int cin_ld) {
  const size_t n = static_cast<size_t>(taps) * cin * cout;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = i % cout, ci = (i / cout) % cin, t = i / (static_cast<size_t>(cout) * cin);
    const __nv_bfloat16 v = __float2bfloat16_rn(W[i]);
    if (wf) wf[(static_cast<size_t>(co) * taps + t) * cin_ld + ci] = v;
    if (wd) wd[(static_cast<size_t>(ci) * taps + (taps - 1 - t)) * cout + co] = v;
  }
}

int pack_conv_weights(const float* W, void* wf, void* wd, int taps, int cin, int cout, int cin_ld, cudaStream_t st) {
  const size_t n = static_cast<size_t>(taps) * cin * cout;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  if (cin_ld < cin) cin_ld = cin;
  pack_conv_weights_kernel<<<grid, 256, 0, st>>>(W, static_cast<__nv_bfloat16*>(wf),
                                                            static_cast<__nv_bfloat16*>(wd), taps, cin, cout, cin_ld);
  DFL_LAUNCH_OK("pack_conv_weights_kernel");
  return DFL_OK;
}

// All 128 -> 128 layers of a generator in ONE launch (blockIdx.y = layer): ptrs = device table [3][n] of
// {W fp32, wf bf16, wd bf16} addresses.  (20 separate launches of ~8 us each were 4 % of the 2D step.)
__global__ void pack_conv_weights_multi_kernel(const unsigned long long* __restrict__ ptrs, int n_layers, int taps, int cin,
                                               int cout) {
  const int l = blockIdx.y;
  const float* __restrict__ W = reinterpret_cast<const float*>(ptrs[l]);
  __nv_bfloat16* __restrict__ wf = reinterpret_cast<__nv_bfloat16*>(ptrs[n_layers + l]);
  __nv_bfloat16* __restrict__ wd = reinterpret_cast<__nv_bfloat16*>(ptrs[2 * n_layers + l]);
  const size_t n = static_cast<size_t>(taps) * cin * cout;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = i % cout, ci = (i / cout) % cin, t = i / (static_cast<size_t>(cout) * cin);
    const __nv_bfloat16 v = __float2bfloat16_rn(W[i]);
    wf[(static_cast<size_t>(co) * taps + t) * cin + ci] = v;
    wd[(static_cast<size_t>(ci) * taps + (taps - 1 - t)) * cout + co] = v;
  }
}

int pack_conv_weights_multi(const void* ptrs, int n_layers, int taps, int cin, int cout, cudaStream_t st) {
  DFL_REQUIRE(ptrs && n_layers >= 1 && n_layers <= 65535, "pack_conv_weights_multi: bad layer table");
  const size_t n = static_cast<size_t>(taps) * cin * cout;
  const int gx = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 2));
  pack_conv_weights_multi_kernel<<<dim3(gx, n_layers), 256, 0, st>>>(static_cast<const unsigned long long*>(ptrs), n_layers,
                                                                    taps, cin, cout);
  DFL_LAUNCH_OK("pack_conv_weights_multi_kernel");
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

// =============================================================================================
// AE / encoder glue (reference model.py:118-216, trainer.py:357-396)
// =============================================================================================
// pad_cast: fp32 [n][cin] -> bf16 [n][128], channels >= cin zero (input of the encoder's first conv)
__global__ void pad_cast_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n, int cin) {
  const size_t total = n * 16;     // 16 threads per voxel, 8 channels each
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int q = idx & 15;
    const size_t v = idx >> 4;
    uint32_t w[4] = {0, 0, 0, 0};
    if (q == 0) {
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < cin && c < 8; ++c) f[c] = in[v * cin + c];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
        w[k] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
    *reinterpret_cast<uint4*>(out + v * 128 + q * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
int pad_cast(const float* in, void* out, size_t n, int cin, cudaStream_t st) {
  DFL_REQUIRE(cin >= 1 && cin <= 8, "pad_cast: 1..8 input channels (got %d)", cin);
  const int grid = static_cast<int>(std::min<size_t>((n * 16 + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  pad_cast_kernel<<<grid, 256, 0, st>>>(in, static_cast<__nv_bfloat16*>(out), n, cin);
  DFL_LAUNCH_OK("pad_cast_kernel");
  return DFL_OK;
}

// add_mask: out = (a + b) * lrelu'(y)   (b and/or y may be null)      bf16 arrays of n elements (n % 8 == 0)
__global__ void add_mask_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out, size_t n8) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 av = __ldg(reinterpret_cast<const uint4*>(a) + i);
    uint4 bv = make_uint4(0, 0, 0, 0), yv = make_uint4(0, 0, 0, 0);
    if (b) bv = __ldg(reinterpret_cast<const uint4*>(b) + i);
    if (y) yv = __ldg(reinterpret_cast<const uint4*>(y) + i);
    const uint32_t aw[4] = {av.x, av.y, av.z, av.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float lo = __uint_as_float(aw[k] << 16) + __uint_as_float(bw[k] << 16);
      float hi = __uint_as_float(aw[k] & 0xFFFF0000u) + __uint_as_float(bw[k] & 0xFFFF0000u);
      if (y) {
        lo *= lrelu_grad_from_out(__uint_as_float(yw[k] << 16));
        hi *= lrelu_grad_from_out(__uint_as_float(yw[k] & 0xFFFF0000u));
      }
      __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
      ow[k] = *reinterpret_cast<uint32_t*>(&h);
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}
int add_mask(const void* a, const void* b, const void* y, void* out, size_t n, cudaStream_t st) {
  DFL_REQUIRE(n % 8 == 0, "add_mask: element count must be a multiple of 8");
  const size_t n8 = n / 8;
  const int grid = static_cast<int>(std::min<size_t>((n8 + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  add_mask_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b),
                                        static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(out), n8);
  DFL_LAUNCH_OK("add_mask_kernel");
  return DFL_OK;
}

// Encoder FC (model.py:149,185): z[b][j] = bias[j] + sum_n flat[b][n] W[n][j], Z <= 16 outputs, B <= 8.
// flat is the channel-blocked concat tensor [nblk][B][V][128] bf16; TF's flatten order is (voxel, channel) with
// channel = blk*128 + c, so n = v*(nblk*128) + blk*128 + c.  One thread per n; block partial sums -> atomics.
constexpr int EF_MAXB = 8, EF_Z = 16;
__global__ void __launch_bounds__(256)
enc_fc_fwd_kernel(const __nv_bfloat16* __restrict__ flat, const float* __restrict__ W, const float* __restrict__ bias,
                  float* __restrict__ z, int B, int V, int nblk, int Z, float* __restrict__ partial) {
  __shared__ float red[8][EF_MAXB * EF_Z];
  const size_t F = static_cast<size_t>(V) * nblk * 128;
  float acc[EF_MAXB][EF_Z];
#pragma unroll
  for (int b = 0; b < EF_MAXB; ++b)
#pragma unroll
    for (int j = 0; j < EF_Z; ++j) acc[b][j] = 0.f;
  for (size_t n = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; n < F;
       n += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = n & 127;
    const int blk = (n >> 7) % nblk;
    const size_t v = (n >> 7) / nblk;
    float w[EF_Z];
#pragma unroll
    for (int j = 0; j < EF_Z; ++j) w[j] = (j < Z) ? __ldg(W + n * Z + j) : 0.f;
#pragma unroll
    for (int b = 0; b < EF_MAXB; ++b) {
      if (b < B) {
        const float f = __bfloat162float(flat[((static_cast<size_t>(blk) * B + b) * V + v) * 128 + c]);
#pragma unroll
        for (int j = 0; j < EF_Z; ++j) acc[b][j] = fmaf(f, w[j], acc[b][j]);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int b = 0; b < EF_MAXB; ++b)
#pragma unroll
    for (int j = 0; j < EF_Z; ++j) {
      const float t = warp_sum(acc[b][j]);
      if (lane == 0) red[warp][b * EF_Z + j] = t;
    }
  __syncthreads();
  if (threadIdx.x < EF_MAXB * EF_Z) {
    const int b = threadIdx.x / EF_Z, j = threadIdx.x % EF_Z;
    if (b < B && j < Z) {
      float t = 0.f;
      for (int w8 = 0; w8 < 8; ++w8) t += red[w8][threadIdx.x];
      if (partial) {                       // deterministic mode: block sums, added in block order by enc_fc_reduce_kernel
        partial[static_cast<size_t>(blockIdx.x) * (EF_MAXB * EF_Z) + threadIdx.x] = t;
      } else {
        if (blockIdx.x == 0) t += bias[j];
        atomicAdd(z + b * Z + j, t);
      }
    }
  }
}
__global__ void __launch_bounds__(EF_MAXB* EF_Z)
enc_fc_reduce_kernel(const float* __restrict__ partial, int nblocks, const float* __restrict__ bias, float* __restrict__ z, int B,
                     int Z) {
  const int b = threadIdx.x / EF_Z, j = threadIdx.x % EF_Z;
  if (b >= B || j >= Z) return;
  float t = bias[j];
  for (int k = 0; k < nblocks; ++k) t += partial[static_cast<size_t>(k) * (EF_MAXB * EF_Z) + threadIdx.x];
  z[b * Z + j] = t;
}
int enc_fc_fwd(const void* flat, const float* W, const float* bias, float* z, int B, int V, int nblk, int Z,
               cudaStream_t st) {
  DFL_REQUIRE(B <= EF_MAXB && Z <= EF_Z, "enc_fc_fwd: B <= %d, Z <= %d (got %d, %d)", EF_MAXB, EF_Z, B, Z);
  DFL_CUDA_OK(cudaMemsetAsync(z, 0, sizeof(float) * B * Z, st));
  const size_t F = static_cast<size_t>(V) * nblk * 128;
  const int grid = static_cast<int>(std::min<size_t>((F + 255) / 256, static_cast<size_t>(num_sms()) * 4));
  size_t wb = 0;
  float* ws = deterministic_workspace(&wb);
  DFL_REQUIRE(!ws || static_cast<size_t>(grid) * EF_MAXB * EF_Z * sizeof(float) <= wb, "enc_fc_fwd: deterministic workspace too small");
  enc_fc_fwd_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(flat), W, bias, z, B, V, nblk, Z, ws);
  DFL_LAUNCH_OK("enc_fc_fwd_kernel");
  if (ws) {
    enc_fc_reduce_kernel<<<1, EF_MAXB * EF_Z, 0, st>>>(ws, grid, bias, z, B, Z);
    DFL_LAUNCH_OK("enc_fc_reduce_kernel");
  }
  return DFL_OK;
}

// backward: dW[n][j] = sum_b flat[b][n] dz[b][j] (written), db[j] = sum_b dz[b][j] (written),
//           dflat[b][n] = sum_j dz[b][j] W[n][j] (bf16, same channel-blocked layout as flat)
__global__ void __launch_bounds__(256)
enc_fc_bwd_kernel(const __nv_bfloat16* __restrict__ flat, const float* __restrict__ W, const float* __restrict__ dz,
                  float* __restrict__ dW, float* __restrict__ db, __nv_bfloat16* __restrict__ dflat, int B, int V,
                  int nblk, int Z) {
  __shared__ float sdz[EF_MAXB * EF_Z];
  if (threadIdx.x < EF_MAXB * EF_Z) {
    const int b = threadIdx.x / EF_Z, j = threadIdx.x % EF_Z;
    sdz[threadIdx.x] = (b < B && j < Z) ? dz[b * Z + j] : 0.f;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < Z) {
    float t = 0.f;
    for (int b = 0; b < B; ++b) t += sdz[b * EF_Z + threadIdx.x];
    db[threadIdx.x] = t;
  }
  const size_t F = static_cast<size_t>(V) * nblk * 128;
  for (size_t n = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; n < F;
       n += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = n & 127;
    const int blk = (n >> 7) % nblk;
    const size_t v = (n >> 7) / nblk;
    float w[EF_Z], g[EF_Z];
#pragma unroll
    for (int j = 0; j < EF_Z; ++j) { w[j] = (j < Z) ? __ldg(W + n * Z + j) : 0.f; g[j] = 0.f; }
#pragma unroll
    for (int b = 0; b < EF_MAXB; ++b) {
      if (b < B) {
        const size_t off = ((static_cast<size_t>(blk) * B + b) * V + v) * 128 + c;
        const float f = __bfloat162float(flat[off]);
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < EF_Z; ++j) {
          g[j] = fmaf(f, sdz[b * EF_Z + j], g[j]);
          d = fmaf(sdz[b * EF_Z + j], w[j], d);
        }
        dflat[off] = __float2bfloat16_rn(d);
      }
    }
#pragma unroll
    for (int j = 0; j < EF_Z; ++j)
      if (j < Z) dW[n * Z + j] = g[j];
  }
}
int enc_fc_bwd(const void* flat, const float* W, const float* dz, float* dW, float* db, void* dflat, int B, int V,
               int nblk, int Z, cudaStream_t st) {
  DFL_REQUIRE(B <= EF_MAXB && Z <= EF_Z, "enc_fc_bwd: B <= %d, Z <= %d (got %d, %d)", EF_MAXB, EF_Z, B, Z);
  const size_t F = static_cast<size_t>(V) * nblk * 128;
  const int grid = static_cast<int>(std::min<size_t>((F + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  enc_fc_bwd_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(flat), W, dz, dW, db,
                                          static_cast<__nv_bfloat16*>(dflat), B, V, nblk, Z);
  DFL_LAUNCH_OK("enc_fc_bwd_kernel");
  return DFL_OK;
}

// decoder-FC input gradient: dz[b][k] (+)= sum_n dout[b][n] W[k][n]     one block per (b, k)
template <typename TI>
__global__ void __launch_bounds__(256)
fc_dz_kernel(const TI* __restrict__ dout, const float* __restrict__ W, float* __restrict__ dz, int K, int N, int accumulate) {
  __shared__ float red[8];
  const int b = blockIdx.x / K, k = blockIdx.x % K;
  float t = 0.f;
  for (int n = threadIdx.x; n < N; n += 256) t = fmaf(ldf(dout + static_cast<size_t>(b) * N + n), __ldg(W + static_cast<size_t>(k) * N + n), t);
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int i = 0; i < 8; ++i) r += red[i];
    dz[b * K + k] = accumulate ? dz[b * K + k] + r : r;
  }
}
int fc_dz(const void* dout, const float* W, float* dz, int B, int K, int N, int dout_dtype, int accumulate, cudaStream_t st) {
  if (dout_dtype == DT_BF16)
    fc_dz_kernel<<<B * K, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dout), W, dz, K, N, accumulate);
  else
    fc_dz_kernel<<<B * K, 256, 0, st>>>(static_cast<const float*>(dout), W, dz, K, N, accumulate);
  DFL_LAUNCH_OK("fc_dz_kernel");
  return DFL_OK;
}

// AE parameter loss (trainer.py:385-387): loss_p = mean((y - z[:, Z-P:])^2);  dz[:, Z-P:] = scale * 2 (z - y) / (B P),
// dz[:, :Z-P] = 0 (dz is then accumulated into by the decoder's fc_dz).   single block
__global__ void ae_loss_p_kernel(const float* __restrict__ z, const float* __restrict__ y, float* __restrict__ dz,
                                 float* __restrict__ loss_p, int B, int Z, int P, float scale) {
  __shared__ float red[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B * Z; i += blockDim.x) {
    const int b = i / Z, j = i % Z;
    float d = 0.f;
    if (j >= Z - P) {
      const float e = z[i] - y[b * P + (j - (Z - P))];
      acc += e * e;
      d = scale * 2.f * e / static_cast<float>(B * P);
    }
    dz[i] = d;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_p = red[0] / static_cast<float>(B * P);
}
int ae_loss_p(const float* z, const float* y, float* dz, float* loss_p, int B, int Z, int P, float scale, cudaStream_t st) {
  DFL_REQUIRE(P >= 1 && P <= Z, "ae_loss_p: need 1 <= p_num <= z_num");
  ae_loss_p_kernel<<<1, 256, 0, st>>>(z, y, dz, loss_p, B, Z, P, scale);
  DFL_LAUNCH_OK("ae_loss_p_kernel");
  return DFL_OK;
}

// use_sparse (config.py:29-30, model.py:196,210, trainer.py:389-394): z = sigmoid(z_lin); loss_kl = sum_j KL(Bernoulli(rho) ||
// Bernoulli(mean_b z[b][j])) over the first Z-P latent dims.  Tiny tensors ([B, Z<=16]): single-block kernels.
__global__ void ae_sigmoid_kernel(const float* __restrict__ zl, float* __restrict__ z, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) z[i] = 1.f / (1.f + __expf(-zl[i]));
}
// dz (gradient w.r.t. the sigmoid OUTPUT, [B,Z]) += w5 * d loss_kl / dz;  then dzl = dz * z (1 - z);  loss_kl written
__global__ void ae_sparse_bwd_kernel(const float* __restrict__ z, float* __restrict__ dz, float* __restrict__ dzl,
                                     float* __restrict__ loss_kl, int B, int Z, int P, float rho, float w5) {
  __shared__ float skl[32];
  const int j = threadIdx.x;
  float kl = 0.f;
  if (j < Z - P) {
    float m = 0.f;
    for (int b = 0; b < B; ++b) m += z[b * Z + j];
    m /= static_cast<float>(B);
    kl = rho * (logf(rho) - logf(m)) + (1.f - rho) * (logf(1.f - rho) - logf(1.f - m));
    const float g = w5 * (-rho / m + (1.f - rho) / (1.f - m)) / static_cast<float>(B);
    for (int b = 0; b < B; ++b) dz[b * Z + j] += g;
  }
  if (j < 32) skl[j] = (j < Z - P) ? kl : 0.f;
  __syncthreads();
  if (j == 0) {
    float t = 0.f;
    for (int k = 0; k < 32; ++k) t += skl[k];
    *loss_kl = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B * Z; i += blockDim.x) dzl[i] = dz[i] * z[i] * (1.f - z[i]);
}
int ae_sigmoid(const float* zl, float* z, int n, cudaStream_t st) {
  ae_sigmoid_kernel<<<1, 128, 0, st>>>(zl, z, n);
  DFL_LAUNCH_OK("ae_sigmoid_kernel");
  return DFL_OK;
}
int ae_sparse_bwd(const float* z, float* dz, float* dzl, float* loss_kl, int B, int Z, int P, float rho, float w5,
                  cudaStream_t st) {
  DFL_REQUIRE(Z <= 32 && P >= 0 && P < Z, "ae_sparse_bwd: need z_num <= 32 and 0 <= p_num < z_num");
  ae_sparse_bwd_kernel<<<1, 64, 0, st>>>(z, dz, dzl, loss_kl, B, Z, P, rho, w5);
  DFL_LAUNCH_OK("ae_sparse_bwd_kernel");
  return DFL_OK;
}

// =============================================================================================
// fp32-grade mode ("bf16x3", BASELINE config 2): every logical fp32 tensor is a (hi, lo) pair of bf16 tensors,
// hi = bf16(v), lo = bf16(v - hi), stored as two consecutive channel blocks.  x*w ~ x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
// (the dropped lo*lo term is 2^-18 relative) is ONE tensor-core GEMM over the virtual input blocks [hi, lo, hi]
// with the weight operand [w_hi | w_hi | w_lo] and fp32 accumulation.
// =============================================================================================
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// W fp32 [taps][cin][cout] -> split operands, K laid out [tap][virtual block 0..2][128 channels] (channels >= cin/cout
// of a padded operand stay zero: zero the buffers once):
//   wf [cout_rows][taps*3*128]:  k = (t*3 + vb)*128 + ci,  vb = 0: w_hi, 1: w_hi, 2: w_lo      (forward)
//   wd [cin_rows ][taps*3*128]:  k = ((taps-1-t)*3 + vb)*128 + co                              (dgrad)
__global__ void pack_split_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ wf,
                                  __nv_bfloat16* __restrict__ wd, int taps, int cin, int cout) {
  const size_t n = static_cast<size_t>(taps) * cin * cout;
  const size_t ld = static_cast<size_t>(taps) * 384;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int co = i % cout, ci = (i / cout) % cin, t = i / (static_cast<size_t>(cout) * cin);
    __nv_bfloat16 hi, lo;
    split_bf16(W[i], hi, lo);
    if (wf) {
      __nv_bfloat16* r = wf + co * ld + static_cast<size_t>(t) * 384 + ci;
      r[0] = hi; r[128] = hi; r[256] = lo;
    }
    if (wd) {
      __nv_bfloat16* r = wd + ci * ld + static_cast<size_t>(taps - 1 - t) * 384 + co;
      r[0] = hi; r[128] = hi; r[256] = lo;
    }
  }
}
int pack_split(const float* W, void* wf, void* wd, int taps, int cin, int cout, cudaStream_t st) {
  DFL_REQUIRE(cin <= 128 && cout <= 128, "pack_split: Cin, Cout <= 128");
  const size_t n = static_cast<size_t>(taps) * cin * cout;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  pack_split_kernel<<<grid, 256, 0, st>>>(W, static_cast<__nv_bfloat16*>(wf), static_cast<__nv_bfloat16*>(wd), taps, cin, cout);
  DFL_LAUNCH_OK("pack_split_kernel");
  return DFL_OK;
}

// fp32 [n][cin] -> (hi, lo) bf16 [2][n][cpad]; channels >= cin zero.  cpad = cin (plain split) or 128 (padded).
__global__ void split_f32_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n, int cin, int cpad) {
  const size_t total = n * cpad;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = idx % cpad;
    const size_t v = idx / cpad;
    __nv_bfloat16 hi = __float2bfloat16_rn(0.f), lo = hi;
    if (c < cin) split_bf16(in[v * cin + c], hi, lo);
    out[idx] = hi;
    out[total + idx] = lo;
  }
}
int split_f32(const float* in, void* out, size_t n, int cin, int cpad, cudaStream_t st) {
  DFL_REQUIRE(cpad >= cin, "split_f32: cpad < cin");
  const int grid = static_cast<int>(std::min<size_t>((n * cpad + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  split_f32_kernel<<<grid, 256, 0, st>>>(in, static_cast<__nv_bfloat16*>(out), n, cin, cpad);
  DFL_LAUNCH_OK("split_f32_kernel");
  return DFL_OK;
}
// (hi, lo) bf16 [2][n] -> fp32 [n]
__global__ void merge_split_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = __bfloat162float(in[i]) + __bfloat162float(in[n + i]);
}
int merge_split(const void* in, float* out, size_t n, cudaStream_t st) {
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  merge_split_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in), out, n);
  DFL_LAUNCH_OK("merge_split_kernel");
  return DFL_OK;
}

// pool_mask on (hi, lo) pairs: ds = sum of the children of g_hi + g_lo; dmasked = ds * lrelu'(mask_hi); outputs split
__global__ void pool_mask_split_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ mask_src,
                                       __nv_bfloat16* __restrict__ ds, __nv_bfloat16* __restrict__ dmasked, int B, int D,
                                       int H, int W, int zr) {
  const size_t ncoarse = static_cast<size_t>(B) * D * H * W;
  const size_t nfine = ncoarse * zr * 4;
  const size_t n = ncoarse * 128;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = idx & 127;
    size_t v = idx >> 7;
    const int x = v % W; v /= W;
    const int y = v % H; v /= H;
    const int z = v % D;
    const int b = v / D;
    const int D2 = D * zr, H2 = 2 * H, W2 = 2 * W;
    float s = 0.f;
    for (int a = 0; a < zr; ++a)
      for (int e = 0; e < 2; ++e)
        for (int f = 0; f < 2; ++f) {
          const size_t pos2 = ((static_cast<size_t>(b) * D2 + (z * zr + a)) * H2 + (2 * y + e)) * W2 + (2 * x + f);
          s += __bfloat162float(g[pos2 * 128 + c]) + __bfloat162float(g[(nfine + pos2) * 128 + c]);
        }
    __nv_bfloat16 hi, lo;
    if (ds) {
      split_bf16(s, hi, lo);
      ds[idx] = hi;
      ds[n + idx] = lo;
    }
    if (dmasked) {
      split_bf16(s * lrelu_grad_from_out(__bfloat162float(mask_src[idx])), hi, lo);
      dmasked[idx] = hi;
      dmasked[n + idx] = lo;
    }
  }
}
int pool_mask_split(const void* g, const void* mask_src, void* ds, void* dmasked, const int64_t* cdims, int nd,
                    cudaStream_t st) {
  const int B = cdims[0], D = nd == 3 ? cdims[1] : 1, H = cdims[nd - 1], W = cdims[nd];
  const size_t n = static_cast<size_t>(B) * D * H * W * 128;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  pool_mask_split_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(g), static_cast<const __nv_bfloat16*>(mask_src),
                                               static_cast<__nv_bfloat16*>(ds), static_cast<__nv_bfloat16*>(dmasked), B, D, H, W,
                                               nd == 3 ? 2 : 1);
  DFL_LAUNCH_OK("pool_mask_split_kernel");
  return DFL_OK;
}


// =============================================================================================
// upscale2 / pool2: nearest x2 up-sampling of a channels-last tensor (ops.py:66-91: out[2i+a] = in[i]) and its adjoint (sum of
// the 2^nd children), for the ops-level API (any channel count, bf16 or fp32).  In the fused engines the up-sampling is the
// conv epilogue's replicated store and the adjoint is pool_mask; these two serve `ops.upscale(3)` as standalone layers.
// =============================================================================================
template <typename U>
__global__ void upscale2_kernel(const U* __restrict__ in, U* __restrict__ out, int B, int D, int H, int W, int units, int nd) {
  // one thread per (fine voxel, unit); units = bytes per voxel / sizeof(U)
  const int zr = nd == 3 ? 2 : 1;
  const size_t n = static_cast<size_t>(B) * (D * zr) * (H * 2) * (W * 2) * units;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t r = i;
    const int u = r % units; r /= units;
    const int x = r % (W * 2); r /= (W * 2);
    const int y = r % (H * 2); r /= (H * 2);
    const int z = r % (D * zr); r /= (D * zr);
    const size_t src = (((r * D + z / zr) * H + (y >> 1)) * W + (x >> 1)) * units + u;
    out[i] = in[src];
  }
}
template <typename T>
__global__ void pool2_kernel(const T* __restrict__ g, T* __restrict__ out, int B, int D, int H, int W, int C, int nd) {
  // one thread per (coarse voxel, channel): fp32 sum of the children in a fixed order
  const int zr = nd == 3 ? 2 : 1;
  const size_t n = static_cast<size_t>(B) * D * H * W * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t r = i;
    const int c = r % C; r /= C;
    const int x = r % W; r /= W;
    const int y = r % H; r /= H;
    const int z = r % D; r /= D;
    float a = 0.f;
    for (int dz = 0; dz < zr; ++dz)
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx)
          a += ldf(g + ((((r * (D * zr) + z * zr + dz) * (H * 2) + 2 * y + dy) * (W * 2)) + 2 * x + dx) * C + c);
    stf(out + i, a);
  }
}
int upscale2(const void* in, void* out, const int64_t* cdims, int nd, int channels, int dtype, cudaStream_t st) {
  DFL_REQUIRE(in && out && (nd == 2 || nd == 3) && channels > 0, "upscale2: null tensor, ndim not 2 / 3 or no channels");
  DFL_REQUIRE(dtype == DFL_F32 || dtype == DFL_BF16, "upscale2: dtype must be DFL_F32 or DFL_BF16");
  const int B = static_cast<int>(cdims[0]), D = nd == 3 ? static_cast<int>(cdims[1]) : 1, H = static_cast<int>(cdims[nd - 1]),
            W = static_cast<int>(cdims[nd]);
  const size_t vb = static_cast<size_t>(channels) * (dtype == DFL_F32 ? 4 : 2);
  const size_t nvox = static_cast<size_t>(B) * D * H * W * (nd == 3 ? 8 : 4);
  if (nvox == 0) return DFL_OK;
  const bool a16 = vb % 16 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const int unit = a16 ? 16 : (vb % 4 == 0 ? 4 : 2);
  const int units = static_cast<int>(vb / unit);
  const int grid = static_cast<int>(std::min<size_t>((nvox * units + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  if (unit == 16) upscale2_kernel<uint4><<<grid, 256, 0, st>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), B, D, H, W, units, nd);
  else if (unit == 4) upscale2_kernel<uint32_t><<<grid, 256, 0, st>>>(static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), B, D, H, W, units, nd);
  else upscale2_kernel<uint16_t><<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), B, D, H, W, units, nd);
  DFL_LAUNCH_OK("upscale2_kernel");
  return DFL_OK;
}
int pool2(const void* g, void* out, const int64_t* cdims, int nd, int channels, int dtype, cudaStream_t st) {
  DFL_REQUIRE(g && out && (nd == 2 || nd == 3) && channels > 0, "pool2: null tensor, ndim not 2 / 3 or no channels");
  DFL_REQUIRE(dtype == DFL_F32 || dtype == DFL_BF16, "pool2: dtype must be DFL_F32 or DFL_BF16");
  const int B = static_cast<int>(cdims[0]), D = nd == 3 ? static_cast<int>(cdims[1]) : 1, H = static_cast<int>(cdims[nd - 1]),
            W = static_cast<int>(cdims[nd]);
  const size_t n = static_cast<size_t>(B) * D * H * W * channels;
  if (n == 0) return DFL_OK;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  if (dtype == DFL_F32) pool2_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(g), static_cast<float*>(out), B, D, H, W, channels, nd);
  else pool2_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(g), static_cast<__nv_bfloat16*>(out), B, D, H, W, channels, nd);
  DFL_LAUNCH_OK("pool2_kernel");
  return DFL_OK;
}

}  // namespace dfl
