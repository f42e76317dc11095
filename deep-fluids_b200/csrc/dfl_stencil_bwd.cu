// deepfluids_b200 -- adjoints of the standalone finite-difference stencils (fp32, one thread per voxel).
//
// The train step never needs these: the fused loss kernels (dfl_stencil.cu, dfl_stencil3_lean.cu) produce dL/dA directly.
// They back the ops-level API, where the reference's ops.curl / ops.jacobian / ops.jacobian3 (ops.py:205-274) are
// differentiable graph nodes -- used e.g. by the discriminator path of arch=dg, whose input is concat(G_, vorticity(G_))
// (trainer.py:149-156, trainer3.py:27-34): TF autodiff of those slice/sub/concat graphs == these kernels.
//
// D = forward difference with the last entry replicated, (D f)[i] = f[i'+1] - f[i'], i' = min(i, n-2).  Adjoint (gather):
//   (D^T g)[k] = gh[k-1] - gh[k],   gh[k] = g[k] (k <= n-3),  g[n-2] + g[n-1] (k = n-2),  0 (k = n-1 or k < 0).
#include "dfl_common.cuh"

namespace dfl {

// (D^T g)[k] along one axis; f(i) returns g at index i of that axis (only called with 0 <= i < n)
template <class F>
__device__ __forceinline__ float adjT(F f, int k, int n) {
  float a = 0.f, b = 0.f;
  if (k - 1 >= 0) a = (k - 1 == n - 2) ? f(k - 1) + f(k) : f(k - 1);      // k-1 <= n-2 always holds for k <= n-1
  if (k < n - 1) b = (k == n - 2) ? f(k) + f(k + 1) : f(k);
  return a - b;
}

// 2D curl (ops.py:264-274): u = D_y psi, v = -D_x psi  =>  dpsi = D_y^T du - D_x^T dv.  dpot has `cs` channels: channel 0
// receives the gradient, the others zero (curl reads channel 0 only).
__global__ void curl2d_bwd_kernel(const float* __restrict__ dvel, float* __restrict__ dpot, int B, int H, int W, int cs) {
  const size_t n = static_cast<size_t>(B) * H * W;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = idx % W, y = (idx / W) % H;
    const size_t base = idx - static_cast<size_t>(y) * W - x;
    const float gy = adjT([&](int i) { return __ldg(dvel + (base + static_cast<size_t>(i) * W + x) * 2); }, y, H);
    const float gx = adjT([&](int i) { return __ldg(dvel + (base + static_cast<size_t>(y) * W + i) * 2 + 1); }, x, W);
    dpot[idx * cs] = gy - gx;
    for (int c = 1; c < cs; ++c) dpot[idx * cs + c] = 0.f;
  }
}

// adjoint of (j, aux) = jacobian(v) / jacobian3(v) w.r.t. v; dj / daux may be null (treated as zero).
// 2D: j = [D_x u, D_y u, D_x v, D_y v], w = D_x v - D_y u
//     du = D_x^T j0 + D_y^T (j1 - w);   dv = D_x^T (j2 + w) + D_y^T j3
// 3D: j = [D_x u, D_y u, D_z u, D_x v, D_y v, D_z v, D_x w, D_y w, D_z w],  c = [D_y w - D_z v, D_z u - D_x w, D_x v - D_y u]
//     du = D_x^T j0 + D_y^T (j1 - c2) + D_z^T (j2 + c1)
//     dv = D_x^T (j3 + c2) + D_y^T j4 + D_z^T (j5 - c0)
//     dw = D_x^T (j6 - c1) + D_y^T (j7 + c0) + D_z^T j8
// The 3D curl alone (trainer3.py:18: `_, G_ = jacobian3(G_s)`) is the dj == null case.
template <int ND>
__global__ void jacobian_bwd_kernel(const float* __restrict__ dj, const float* __restrict__ daux, float* __restrict__ dv,
                                    int B, int D, int H, int W) {
  constexpr int NJ = ND * ND, NA = ND == 2 ? 1 : 3;
  const size_t n = static_cast<size_t>(B) * D * H * W;
  const size_t hw = static_cast<size_t>(H) * W;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = idx % W, y = (idx / W) % H, z = ND == 3 ? static_cast<int>((idx / hw) % D) : 0;
    const size_t base = idx - (static_cast<size_t>(z) * H + y) * W - x;
    // field value (component comp, axis a) at position with coordinate i substituted along axis a
    auto pos = [&](int a, int i) -> size_t {
      return base + (static_cast<size_t>(a == 2 ? i : z) * H + (a == 1 ? i : y)) * W + (a == 0 ? i : x);
    };
    // g(comp, axis, position): the upstream gradient of D_axis v_comp, i.e. dj entry + signed aux entries
    auto g = [&](int comp, int a, size_t p) -> float {
      float v = dj ? __ldg(dj + p * NJ + comp * ND + a) : 0.f;
      if (daux) {
        if (ND == 2) {
          if (comp == 1 && a == 0) v += __ldg(daux + p);            // w = +D_x v
          if (comp == 0 && a == 1) v -= __ldg(daux + p);            //     -D_y u
        } else {
          // c0 = D_y w - D_z v;  c1 = D_z u - D_x w;  c2 = D_x v - D_y u
          if (comp == 2 && a == 1) v += __ldg(daux + p * NA + 0);
          if (comp == 1 && a == 2) v -= __ldg(daux + p * NA + 0);
          if (comp == 0 && a == 2) v += __ldg(daux + p * NA + 1);
          if (comp == 2 && a == 0) v -= __ldg(daux + p * NA + 1);
          if (comp == 1 && a == 0) v += __ldg(daux + p * NA + 2);
          if (comp == 0 && a == 1) v -= __ldg(daux + p * NA + 2);
        }
      }
      return v;
    };
    const int coord[3] = {x, y, z}, ext[3] = {W, H, D};
#pragma unroll
    for (int comp = 0; comp < ND; ++comp) {
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < ND; ++a) acc += adjT([&](int i) { return g(comp, a, pos(a, i)); }, coord[a], ext[a]);
      dv[idx * ND + comp] = acc;
    }
  }
}

// LSGAN terms of arch=dg (trainer.py:174-176 / trainer3.py:53-55): loss = mean((d - target)^2) over the patch
// discriminator's output; dd = scale * d loss / d d (optional).  One block, fixed summation order (deterministic).
__global__ void __launch_bounds__(256) mse_loss_kernel(const float* __restrict__ d, float target, float* __restrict__ loss,
                                                       float* __restrict__ dd, size_t n, float scale) {
  __shared__ double sred[8];
  double acc = 0.0;
  const float k = 2.f * scale / static_cast<float>(n);
  for (size_t i = threadIdx.x; i < n; i += 256) {
    const float e = d[i] - target;
    acc += static_cast<double>(e) * e;
    if (dd) dd[i] = k * e;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += sred[i];
    *loss = static_cast<float>(t / static_cast<double>(n));
  }
}

int mse_loss(const float* d, float target, float* loss, float* dd, size_t n, float scale, cudaStream_t st) {
  DFL_REQUIRE(d && loss && n > 0, "mse_loss: null tensor or empty");
  mse_loss_kernel<<<1, 256, 0, st>>>(d, target, loss, dd, n, scale);
  DFL_LAUNCH_OK("mse_loss_kernel");
  return DFL_OK;
}

// L1 term of the loss when it is NOT fused with the curl (use_curl=False, trainer.py:141-144,170-171): loss = mean|a - b|,
// dd (+)= scale * sgn(a - b) / n.  Per-block fp64 partials, summed in block order by the last block to finish (atomic
// ticket) -> deterministic.  workspace: (L1_MAX_BLOCKS + 1) doubles, zeroed by the caller once (the ticket resets itself).
constexpr int L1_MAX_BLOCKS = 1024;
__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      float* __restrict__ loss, float* __restrict__ dd, size_t n, float scale,
                                                      int accumulate, double* __restrict__ ws) {
  __shared__ double sred[8];
  __shared__ bool last;
  double acc = 0.0;
  const float k = scale / static_cast<float>(n);
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += 256ull * gridDim.x) {
    const float e = a[i] - b[i];
    acc += fabsf(e);
    if (dd) {
      const float g = k * ((e > 0.f ? 1.f : 0.f) - (e < 0.f ? 1.f : 0.f));
      dd[i] = accumulate ? dd[i] + g : g;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += sred[i];
    ws[1 + blockIdx.x] = t;
    __threadfence();
    unsigned long long* ticket = reinterpret_cast<unsigned long long*>(ws);
    last = (atomicAdd(ticket, 1ull) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double t = 0;
    for (unsigned i = 0; i < gridDim.x; ++i) t += reinterpret_cast<volatile double*>(ws)[1 + i];
    *loss = static_cast<float>(t / static_cast<double>(n));
    *reinterpret_cast<unsigned long long*>(ws) = 0ull;
  }
}

size_t l1_loss_workspace_bytes() { return (L1_MAX_BLOCKS + 1) * sizeof(double); }

int l1_loss(const float* a, const float* b, float* loss, float* dd, size_t n, float scale, int accumulate, void* ws,
            cudaStream_t st) {
  DFL_REQUIRE(a && b && loss && ws && n > 0, "l1_loss: null tensor or empty");
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, std::min<size_t>(L1_MAX_BLOCKS, static_cast<size_t>(num_sms()) * 4)));
  l1_loss_kernel<<<grid, 256, 0, st>>>(a, b, loss, dd, n, scale, accumulate, static_cast<double*>(ws));
  DFL_LAUNCH_OK("l1_loss_kernel");
  return DFL_OK;
}

// op 0: dpot = curl^T(dvel) (2D: dpot_cs channels, 3D: 3);  op 1: dvel = jacobian^T(dj, daux)
int bwd_stencils(int op, int nd, const int64_t* dims, const float* g0, const float* g1, float* out, int out_cs,
                 cudaStream_t st) {
  DFL_REQUIRE(nd == 2 || nd == 3, "stencil_bwd: ndim must be 2 or 3 (got %d)", nd);
  for (int k = 1; k <= nd; ++k) DFL_REQUIRE(dims[k] >= 2, "stencil_bwd: every spatial extent must be >= 2");
  DFL_REQUIRE(out != nullptr, "stencil_bwd: null output");
  const int B = static_cast<int>(dims[0]);
  const int D = nd == 3 ? static_cast<int>(dims[1]) : 1, H = static_cast<int>(dims[nd - 1]), W = static_cast<int>(dims[nd]);
  const size_t n = static_cast<size_t>(B) * D * H * W;
  const int threads = 256;
  const int grid = static_cast<int>(std::min<size_t>((n + threads - 1) / threads, static_cast<size_t>(num_sms()) * 16));
  if (op == 0) {
    DFL_REQUIRE(g0 != nullptr, "curl_bwd: null upstream gradient");
    if (nd == 2) {
      DFL_REQUIRE(out_cs >= 1, "curl_bwd: dpot needs >= 1 channel");
      curl2d_bwd_kernel<<<grid, threads, 0, st>>>(g0, out, B, H, W, out_cs);
    } else {
      DFL_REQUIRE(out_cs == 3, "curl_bwd (3D): the potential has exactly 3 channels");
      jacobian_bwd_kernel<3><<<grid, threads, 0, st>>>(nullptr, g0, out, B, D, H, W);     // 3D curl = jacobian3's second output
    }
  } else if (op == 1) {
    DFL_REQUIRE(g0 != nullptr || g1 != nullptr, "jacobian_bwd: both upstream gradients are null");
    if (nd == 2)
      jacobian_bwd_kernel<2><<<grid, threads, 0, st>>>(g0, g1, out, B, 1, H, W);
    else
      jacobian_bwd_kernel<3><<<grid, threads, 0, st>>>(g0, g1, out, B, D, H, W);
  } else {
    set_last_error("unknown stencil_bwd op %d", op);
    return DFL_ERR_ARG;
  }
  DFL_LAUNCH_OK("stencil_bwd_kernel");
  return DFL_OK;
}

}  // namespace dfl
