// deepfluids_b200 -- finite-difference stencil kernels (HBM-bound part of the hot path).
//
// Replaces, for the generator/AE train step, the ~150-200 tiny TF slice/sub/concat ops of
//   reference ops.py:264-274 (curl), ops.py:205-262 (jacobian / jacobian3, incl. the 3D curl),
//   trainer.py:170-172 / trainer3.py:49-51 (L1 + Jacobian-L1 loss) and their autodiff adjoints
// by ONE fused pass:   (potential A, target x)  ->  (loss sums, dL/dA [, G_ = curl(A)]).
//
// Math.  D_a = forward difference along axis a with the last entry replicated:
//   (D f)[i] = f[i'+1] - f[i'],  i' = min(i, n-2).
// Its adjoint is a plain backward difference of the "folded" field:
//   (D^T g)[k] = gh[k-1] - gh[k],  gh[k] = g[k] (k<=n-3), g[n-2]+g[n-1] (k=n-2), 0 (k=n-1 or k<0).
// With G = curl(A), e = G - x:
//   L = w1*mean|e| + w2*mean|D_a G_c - D_a x_c|      (means over all elements, trainer.py:170-171)
//   dL/dG_c[p] = c1*sgn(e_c[p]) + c2*sum_a ( sh_{c,a}[p-1_a] - sh_{c,a}[p] ),
//   sh_{c,a}[q] = wgt(q_a) * sgn( (G_c[q+1_a]-G_c[q]) - (x_c[q+1_a]-x_c[q]) ),  wgt = 1, 2 at n-2, 0 at n-1
//   dL/dA = curl^T(dL/dG).
// The differences are evaluated in the reference's own order so G_ is bit-identical to ops.curl /
// ops.jacobian3 in fp32; only the order of the loss summation differs.
//
// 3D kernel: one thread per (x,y) column of a halo'd tile, marching along z with a 3-stage software
// pipeline (G, dL/dG, dL/dA) whose in-plane neighbours are staged through double-buffered shared-memory
// planes and whose z neighbours live in registers: every A/x plane is read once per tile.
#include "dfl_common.cuh"

namespace dfl {

__device__ __forceinline__ float sgnf(float v) { return (v > 0.f ? 1.f : 0.f) - (v < 0.f ? 1.f : 0.f); }
__device__ __forceinline__ float wgt(int k, int n) { return (k < 0 || k >= n - 1) ? 0.f : (k == n - 2 ? 2.f : 1.f); }
// folded value gh[k] given g[k] and g[k+1]
__device__ __forceinline__ float foldv(float g0, float g1, int k, int n) {
  return (k < 0 || k >= n - 1) ? 0.f : (k == n - 2 ? g0 + g1 : g0);
}

struct StencilParams {
  int B, D, H, W;       // D == 1 for 2D
  int pot_cs;           // channel stride (number of channels) of the potential tensor
  int dpot_cs;          // channel stride of the gradient tensor (2D: 1, or pot_cs when the caller wants a full-shape gradient)
  float c1, c2;         // w1/N1 * grad_scale, w2/N2 * grad_scale
  int zseg;             // z planes per block (3D)
  int tiles_x, tiles_y, nseg;
};

// =============================================================================================
// 3D fused kernel
// =============================================================================================
constexpr int S3_TX = 32, S3_TY = 16;          // threads = halo'd tile columns
constexpr int S3_OX = S3_TX - 5, S3_OY = S3_TY - 5;  // output columns per tile

template <typename TP, typename TX_, typename TO, bool kWriteVel>
__global__ void __launch_bounds__(S3_TX* S3_TY, 2)
stencil3d_fused_kernel(const TP* __restrict__ pot, const TX_* __restrict__ xt, TO* __restrict__ dpot,
                       TO* __restrict__ vel, double* __restrict__ partials, StencilParams p) {
  __shared__ __align__(16) float sA[2][S3_TY][S3_TX][3];
  __shared__ float sG[2][S3_TY][S3_TX][3];
  __shared__ float sX[2][S3_TY][S3_TX][3];
  __shared__ float sD[2][S3_TY][S3_TX][3];
  // the block-reduction scratch aliases sA (free after the last iteration's barrier)
  double (*sred)[S3_TX * S3_TY / 32] = reinterpret_cast<double (*)[S3_TX * S3_TY / 32]>(&sA[0][0][0][0]);

  const int i = threadIdx.x, j = threadIdx.y;
  int blk = blockIdx.x;
  const int tx = blk % p.tiles_x;
  blk /= p.tiles_x;
  const int ty = blk % p.tiles_y;
  blk /= p.tiles_y;
  const int seg = blk % p.nseg;
  const int b = blk / p.nseg;

  const int D = p.D, H = p.H, W = p.W;
  const int cx = tx * S3_OX - 2 + i, cy = ty * S3_OY - 2 + j;
  const bool in_xy = (cx >= 0 && cx < W && cy >= 0 && cy < H);
  const bool out_col = in_xy && i >= 2 && i < S3_TX - 3 && j >= 2 && j < S3_TY - 3;
  const int zs = seg * p.zseg, ze = min(D, zs + p.zseg);
  const size_t col = (static_cast<size_t>(b) * D * H + (in_xy ? cy : 0)) * W + (in_xy ? cx : 0);
  const size_t plane = static_cast<size_t>(H) * W;

  // clamped in-plane neighbour indices inside the block's smem planes
  const int im = max(i - 1, 0), ip = min(i + 1, S3_TX - 1);
  const int jm = max(j - 1, 0), jp = min(j + 1, S3_TY - 1);
  // replicate-last selection for the in-plane differences of plane values: (lo, hi) index pair
  const int ix_lo = (cx <= W - 2) ? i : im, ix_hi = (cx <= W - 2) ? ip : i;
  const int jy_lo = (cy <= H - 2) ? j : jm, jy_hi = (cy <= H - 2) ? jp : j;

  float a_prev[3] = {0, 0, 0}, a_cur[3] = {0, 0, 0}, a_nxt[3];
  float g_prev[3] = {0, 0, 0}, g_cur[3] = {0, 0, 0}, g_new[3];
  float x_prev[3] = {0, 0, 0}, x_cur[3] = {0, 0, 0}, x_new[3];
  float d_prev[3] = {0, 0, 0}, d_cur[3] = {0, 0, 0}, d_new[3];
  float facc_l1 = 0.f, facc_j = 0.f;  // <= zseg*12 terms per thread: fp32 is ample, fp64 only across threads
  float sz_prev[3] = {0, 0, 0};       // wz0*sgn(dzp) of the previous plane == the backward z term of this plane

  // software prefetch: A[t+1] and x[t] are requested one iteration before they are consumed, so the global-load
  // latency overlaps the previous iteration's arithmetic and barrier
  float a_pf[3] = {0, 0, 0}, x_pf[3] = {0, 0, 0};
  {
    const int t0 = zs - 2;
    if (t0 >= 0 && t0 < D && in_xy) {
      const TP* q = pot + (col + static_cast<size_t>(t0) * plane) * p.pot_cs;
      a_pf[0] = ldf(q); a_pf[1] = ldf(q + 1); a_pf[2] = ldf(q + 2);
    }
    const int q0 = t0 - 1;
    if (q0 >= 0 && q0 < D && in_xy) {
      const TX_* xp = xt + (col + static_cast<size_t>(q0) * plane) * 3;
      x_pf[0] = ldf(xp); x_pf[1] = ldf(xp + 1); x_pf[2] = ldf(xp + 2);
    }
  }

  for (int t = zs - 2; t <= ze + 2; ++t) {
    const int cur = t & 1, prv = cur ^ 1;
    // ---- A[t], x[t-1] were prefetched; request A[t+1], x[t] now ----
    a_nxt[0] = a_pf[0]; a_nxt[1] = a_pf[1]; a_nxt[2] = a_pf[2];
    x_new[0] = x_pf[0]; x_new[1] = x_pf[1]; x_new[2] = x_pf[2];
    a_pf[0] = a_pf[1] = a_pf[2] = 0.f;
    x_pf[0] = x_pf[1] = x_pf[2] = 0.f;
    if (in_xy && t + 1 <= ze + 2) {
      if (t + 1 >= 0 && t + 1 < D) {
        const TP* q = pot + (col + static_cast<size_t>(t + 1) * plane) * p.pot_cs;
        a_pf[0] = ldf(q); a_pf[1] = ldf(q + 1); a_pf[2] = ldf(q + 2);
      }
      if (t >= 0 && t < D) {
        const TX_* xp = xt + (col + static_cast<size_t>(t) * plane) * 3;
        x_pf[0] = ldf(xp); x_pf[1] = ldf(xp + 1); x_pf[2] = ldf(xp + 2);
      }
    }
    sA[cur][j][i][0] = a_nxt[0];
    sA[cur][j][i][1] = a_nxt[1];
    sA[cur][j][i][2] = a_nxt[2];

    // ---- stage G: q = t-1  (in-plane neighbours of A[q] were staged last iteration) ----
    {
      const int q = t - 1;
      const bool qin = (q >= 0 && q < D) && in_xy;
      if (qin) {
        // z differences (replicate last): q <= D-2 ? A[q+1]-A[q] : A[q]-A[q-1]
        const bool zl = (q <= D - 2);
        const float dudz = zl ? (a_nxt[0] - a_cur[0]) : (a_cur[0] - a_prev[0]);
        const float dvdz = zl ? (a_nxt[1] - a_cur[1]) : (a_cur[1] - a_prev[1]);
        const float dwdy = sA[prv][jy_hi][i][2] - sA[prv][jy_lo][i][2];
        const float dudy = sA[prv][jy_hi][i][0] - sA[prv][jy_lo][i][0];
        const float dwdx = sA[prv][j][ix_hi][2] - sA[prv][j][ix_lo][2];
        const float dvdx = sA[prv][j][ix_hi][1] - sA[prv][j][ix_lo][1];
        g_new[0] = dwdy - dvdz;   // ops.py:255
        g_new[1] = dudz - dwdx;   // ops.py:256
        g_new[2] = dvdx - dudy;   // ops.py:257
        if (kWriteVel && out_col && q >= zs && q < ze) {
          TO* vp = vel + (col + static_cast<size_t>(q) * plane) * 3;
          stf(vp, g_new[0]);
          stf(vp + 1, g_new[1]);
          stf(vp + 2, g_new[2]);
        }
      } else {
        g_new[0] = g_new[1] = g_new[2] = 0.f;
        x_new[0] = x_new[1] = x_new[2] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        sG[cur][j][i][c] = g_new[c];
        sX[cur][j][i][c] = x_new[c];
      }
    }

    // ---- stage dG: q = t-2  (in-plane neighbours of G[q], x[q] were staged last iteration) ----
    {
      const int q = t - 2;
      const bool qin = (q >= 0 && q < D) && in_xy;
      const float wxm = wgt(cx - 1, W), wx0 = wgt(cx, W);
      const float wym = wgt(cy - 1, H), wy0 = wgt(cy, H);
      const float wz0 = wgt(q, D);
      const bool count = out_col && q >= zs && q < ze;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float g0 = g_cur[c], x0 = x_cur[c];
        const float e = g0 - x0;
        // x axis
        const float dxp = (sG[prv][j][ip][c] - g0) - (sX[prv][j][ip][c] - x0);
        const float dxm = (g0 - sG[prv][j][im][c]) - (x0 - sX[prv][j][im][c]);
        const float dyp = (sG[prv][jp][i][c] - g0) - (sX[prv][jp][i][c] - x0);
        const float dym = (g0 - sG[prv][jm][i][c]) - (x0 - sX[prv][jm][i][c]);
        const float dzp = (g_new[c] - g0) - (x_new[c] - x0);
        const float szp = wz0 * sgnf(dzp);      // the forward z term of plane q is the backward term of plane q+1
        float dg = p.c1 * sgnf(e) +
                   p.c2 * ((wxm * sgnf(dxm) - wx0 * sgnf(dxp)) + (wym * sgnf(dym) - wy0 * sgnf(dyp)) +
                           (sz_prev[c] - szp));
        sz_prev[c] = qin ? szp : 0.f;
        d_new[c] = qin ? dg : 0.f;
        if (count) {
          facc_l1 += fabsf(e);
          facc_j += wx0 * fabsf(dxp) + wy0 * fabsf(dyp) + wz0 * fabsf(dzp);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) sD[cur][j][i][c] = d_new[c];
    }

    // ---- stage dA: r = t-3  (in-plane neighbours of dG[r] were staged last iteration) ----
    {
      const int r = t - 3;
      if (out_col && r >= zs && r < ze) {
        // gU=0, gV=1, gW=2
        // D_z^T g [r] = gh[r-1] - gh[r]
        const float dzT_V = foldv(d_prev[1], d_cur[1], r - 1, D) - foldv(d_cur[1], d_new[1], r, D);
        const float dzT_U = foldv(d_prev[0], d_cur[0], r - 1, D) - foldv(d_cur[0], d_new[0], r, D);
        const float dyT_W = foldv(sD[prv][jm][i][2], d_cur[2], cy - 1, H) - foldv(d_cur[2], sD[prv][jp][i][2], cy, H);
        const float dyT_U = foldv(sD[prv][jm][i][0], d_cur[0], cy - 1, H) - foldv(d_cur[0], sD[prv][jp][i][0], cy, H);
        const float dxT_W = foldv(sD[prv][j][im][2], d_cur[2], cx - 1, W) - foldv(d_cur[2], sD[prv][j][ip][2], cx, W);
        const float dxT_V = foldv(sD[prv][j][im][1], d_cur[1], cx - 1, W) - foldv(d_cur[1], sD[prv][j][ip][1], cx, W);
        TO* o = dpot + (col + static_cast<size_t>(r) * plane) * 3;
        stf(o, dzT_V - dyT_W);      // dA_u = D_z^T gV - D_y^T gW
        stf(o + 1, dxT_W - dzT_U);  // dA_v = D_x^T gW - D_z^T gU
        stf(o + 2, dyT_U - dxT_V);  // dA_w = D_y^T gU - D_x^T gV
      }
    }

    // ---- rotate registers ----
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      a_prev[c] = a_cur[c]; a_cur[c] = a_nxt[c];
      g_prev[c] = g_cur[c]; g_cur[c] = g_new[c];
      x_prev[c] = x_cur[c]; x_cur[c] = x_new[c];
      d_prev[c] = d_cur[c]; d_cur[c] = d_new[c];
    }
    __syncthreads();
  }

  // ---- block reduction of the two loss sums -> partials[block] ----
  double acc_l1 = warp_sum(static_cast<double>(facc_l1));
  double acc_j = warp_sum(static_cast<double>(facc_j));
  const int tid = j * S3_TX + i, wid = tid >> 5;
  if ((tid & 31) == 0) {
    sred[0][wid] = acc_l1;
    sred[1][wid] = acc_j;
  }
  __syncthreads();
  if (tid == 0) {
    double a = 0, c = 0;
    for (int k = 0; k < S3_TX * S3_TY / 32; ++k) {
      a += sred[0][k];
      c += sred[1][k];
    }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = c;
  }
}

// NOTE on the stage-dG register usage: at the time dG[q=t-2] is formed, g_cur == G[t-2], g_prev == G[t-3],
// g_new == G[t-1] (just computed) -- the rotation happens at the end of the iteration.  Likewise for dA[r=t-3]:
// d_cur == dG[t-3], d_prev == dG[t-4], d_new == dG[t-2].

// =============================================================================================
// 2D fused kernel: one pixel per thread, halo'd tile fully in smem, three block-synchronised stages
// =============================================================================================
constexpr int S2_TX = 32, S2_TY = 32;
constexpr int S2_OX = S2_TX - 5, S2_OY = S2_TY - 5;

template <typename TP, typename TX_, typename TO, bool kWriteVel>
__global__ void __launch_bounds__(S2_TX* S2_TY)
stencil2d_fused_kernel(const TP* __restrict__ pot, const TX_* __restrict__ xt, TO* __restrict__ dpot,
                       TO* __restrict__ vel, double* __restrict__ partials, StencilParams p) {
  __shared__ float sP[S2_TY][S2_TX + 1];
  __shared__ float sG[S2_TY][S2_TX + 1][2];
  __shared__ float sX[S2_TY][S2_TX + 1][2];
  __shared__ float sD[S2_TY][S2_TX + 1][2];
  __shared__ double sred[2][S2_TX * S2_TY / 32];

  const int i = threadIdx.x, j = threadIdx.y;
  int blk = blockIdx.x;
  const int tx = blk % p.tiles_x;
  blk /= p.tiles_x;
  const int ty = blk % p.tiles_y;
  const int b = blk / p.tiles_y;
  const int H = p.H, W = p.W;
  const int cx = tx * S2_OX - 2 + i, cy = ty * S2_OY - 2 + j;
  const bool in_xy = (cx >= 0 && cx < W && cy >= 0 && cy < H);
  const bool out_px = in_xy && i >= 2 && i < S2_TX - 3 && j >= 2 && j < S2_TY - 3;
  const size_t pix = (static_cast<size_t>(b) * H + (in_xy ? cy : 0)) * W + (in_xy ? cx : 0);
  const int im = max(i - 1, 0), ip = min(i + 1, S2_TX - 1);
  const int jm = max(j - 1, 0), jp = min(j + 1, S2_TY - 1);
  const int ix_lo = (cx <= W - 2) ? i : im, ix_hi = (cx <= W - 2) ? ip : i;
  const int jy_lo = (cy <= H - 2) ? j : jm, jy_hi = (cy <= H - 2) ? jp : j;

  float x0 = 0.f, x1 = 0.f;
  sP[j][i] = in_xy ? ldf(pot + pix * p.pot_cs) : 0.f;
  if (in_xy) {
    x0 = ldf(xt + pix * 2);
    x1 = ldf(xt + pix * 2 + 1);
  }
  sX[j][i][0] = x0;
  sX[j][i][1] = x1;
  __syncthreads();
  // G = curl(psi): u = d psi/dy, v = psi[x] - psi[x+1]   (ops.py:267-270)
  float g0 = 0.f, g1 = 0.f;
  if (in_xy) {
    g0 = sP[jy_hi][i] - sP[jy_lo][i];
    g1 = sP[j][ix_lo] - sP[j][ix_hi];
    if (kWriteVel && out_px) {
      stf(vel + pix * 2, g0);
      stf(vel + pix * 2 + 1, g1);
    }
  }
  sG[j][i][0] = g0;
  sG[j][i][1] = g1;
  __syncthreads();
  double acc_l1 = 0.0, acc_j = 0.0;
  {
    const float wxm = wgt(cx - 1, W), wx0 = wgt(cx, W);
    const float wym = wgt(cy - 1, H), wy0 = wgt(cy, H);
    const float gg[2] = {g0, g1}, xx[2] = {x0, x1};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float e = gg[c] - xx[c];
      const float dxp = (sG[j][ip][c] - gg[c]) - (sX[j][ip][c] - xx[c]);
      const float dxm = (gg[c] - sG[j][im][c]) - (xx[c] - sX[j][im][c]);
      const float dyp = (sG[jp][i][c] - gg[c]) - (sX[jp][i][c] - xx[c]);
      const float dym = (gg[c] - sG[jm][i][c]) - (xx[c] - sX[jm][i][c]);
      const float dg = p.c1 * sgnf(e) +
                       p.c2 * ((wxm * sgnf(dxm) - wx0 * sgnf(dxp)) + (wym * sgnf(dym) - wy0 * sgnf(dyp)));
      sD[j][i][c] = in_xy ? dg : 0.f;
      if (out_px) {
        acc_l1 += fabsf(e);
        acc_j += static_cast<double>(wx0 * fabsf(dxp) + wy0 * fabsf(dyp));
      }
    }
  }
  __syncthreads();
  if (out_px) {
    // d psi = D_y^T gU - D_x^T gV'   with v = -(D_x psi)  =>  d psi = D_y^T gU - D_x^T gV
    const float dyT_U = foldv(sD[jm][i][0], sD[j][i][0], cy - 1, H) - foldv(sD[j][i][0], sD[jp][i][0], cy, H);
    const float dxT_V = foldv(sD[j][im][1], sD[j][i][1], cx - 1, W) - foldv(sD[j][i][1], sD[j][ip][1], cx, W);
    stf(dpot + pix * p.dpot_cs, dyT_U - dxT_V);
    for (int c = 1; c < p.dpot_cs; ++c) stf(dpot + pix * p.dpot_cs + c, 0.f);   // unused output channels (2D AE)
  }
  acc_l1 = warp_sum(acc_l1);
  acc_j = warp_sum(acc_j);
  const int tid = j * S2_TX + i, wid = tid >> 5;
  if ((tid & 31) == 0) {
    sred[0][wid] = acc_l1;
    sred[1][wid] = acc_j;
  }
  __syncthreads();
  if (tid == 0) {
    double a = 0, c = 0;
    for (int k = 0; k < S2_TX * S2_TY / 32; ++k) {
      a += sred[0][k];
      c += sred[1][k];
    }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = c;
  }
}

// loss3 = {total, l1, j_l1}; fixed-order (deterministic) reduction of the per-block partial sums
__global__ void stencil_finalize_kernel(const double* __restrict__ partials, int nblk, double inv_n1, double inv_n2,
                                        float w1, float w2, float* __restrict__ loss3) {
  __shared__ double s0[256], s1[256];
  double a = 0, c = 0;
  for (int k = threadIdx.x; k < nblk; k += 256) {
    a += partials[2 * k];
    c += partials[2 * k + 1];
  }
  s0[threadIdx.x] = a;
  s1[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s0[threadIdx.x] += s0[threadIdx.x + o];
      s1[threadIdx.x] += s1[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double l1 = s0[0] * inv_n1, jl = s1[0] * inv_n2;
    loss3[0] = static_cast<float>(w1 * l1 + w2 * jl);
    loss3[1] = static_cast<float>(l1);
    loss3[2] = static_cast<float>(jl);
  }
}

// =============================================================================================
// standalone forward stencils (ops.curl / ops.jacobian / ops.jacobian3 / ops.divergence*), one thread per voxel.
// These back the ops-level API; the train step uses the fused kernels above.
// =============================================================================================
template <typename T>
__global__ void curl2d_kernel(const T* __restrict__ pot, T* __restrict__ vel, int B, int H, int W, int cs) {
  const size_t n = static_cast<size_t>(B) * H * W;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = idx % W, y = (idx / W) % H;
    const int yl = min(y, H - 2), xl = min(x, W - 2);
    const size_t base = idx - static_cast<size_t>(y) * W - x;
    const float u = ldf(pot + (base + static_cast<size_t>(yl + 1) * W + x) * cs) -
                    ldf(pot + (base + static_cast<size_t>(yl) * W + x) * cs);
    const float v = ldf(pot + (base + static_cast<size_t>(y) * W + xl) * cs) -
                    ldf(pot + (base + static_cast<size_t>(y) * W + xl + 1) * cs);
    stf(vel + idx * 2, u);
    stf(vel + idx * 2 + 1, v);
  }
}

// jacobian (2D): j = [dudx,dudy,dvdx,dvdy], w = dvdx - dudy
template <typename T>
__global__ void jacobian2d_kernel(const T* __restrict__ v, T* __restrict__ jac, T* __restrict__ vort, int B, int H,
                                  int W) {
  const size_t n = static_cast<size_t>(B) * H * W;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = idx % W, y = (idx / W) % H;
    const int yl = min(y, H - 2), xl = min(x, W - 2);
    const size_t base = idx - static_cast<size_t>(y) * W - x;
    const size_t px0 = base + static_cast<size_t>(y) * W + xl, py0 = base + static_cast<size_t>(yl) * W + x;
    float d[4];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      d[2 * c] = ldf(v + (px0 + 1) * 2 + c) - ldf(v + px0 * 2 + c);
      d[2 * c + 1] = ldf(v + (py0 + W) * 2 + c) - ldf(v + py0 * 2 + c);
    }
    if (jac) {
#pragma unroll
      for (int k = 0; k < 4; ++k) stf(jac + idx * 4 + k, d[k]);
    }
    if (vort) stf(vort + idx, d[2] - d[1]);
  }
}

// jacobian3: j = [dudx,dudy,dudz,dvdx,dvdy,dvdz,dwdx,dwdy,dwdz], c = [dwdy-dvdz, dudz-dwdx, dvdx-dudy]
template <typename T>
__global__ void jacobian3d_kernel(const T* __restrict__ v, T* __restrict__ jac, T* __restrict__ curl, int B, int D,
                                  int H, int W) {
  const size_t n = static_cast<size_t>(B) * D * H * W;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = idx % W, y = (idx / W) % H, z = (idx / (static_cast<size_t>(W) * H)) % D;
    const int xl = min(x, W - 2), yl = min(y, H - 2), zl = min(z, D - 2);
    const size_t hw = static_cast<size_t>(H) * W;
    const size_t base = idx - (static_cast<size_t>(z) * H + y) * W - x;
    const size_t px0 = base + (static_cast<size_t>(z) * H + y) * W + xl;
    const size_t py0 = base + (static_cast<size_t>(z) * H + yl) * W + x;
    const size_t pz0 = base + (static_cast<size_t>(zl) * H + y) * W + x;
    float d[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      d[3 * c] = ldf(v + (px0 + 1) * 3 + c) - ldf(v + px0 * 3 + c);
      d[3 * c + 1] = ldf(v + (py0 + W) * 3 + c) - ldf(v + py0 * 3 + c);
      d[3 * c + 2] = ldf(v + (pz0 + hw) * 3 + c) - ldf(v + pz0 * 3 + c);
    }
    if (jac) {
#pragma unroll
      for (int k = 0; k < 9; ++k) stf(jac + idx * 9 + k, d[k]);
    }
    if (curl) {
      stf(curl + idx * 3, d[7] - d[5]);
      stf(curl + idx * 3 + 1, d[2] - d[6]);
      stf(curl + idx * 3 + 2, d[3] - d[1]);
    }
  }
}

// divergence on the [:-1] interior (ops.py:276-290): out [B,(D-1,)H-1,W-1,1]
template <typename T>
__global__ void divergence_kernel(const T* __restrict__ v, T* __restrict__ out, int B, int D, int H, int W, int nd) {
  const int Do = (nd == 3) ? D - 1 : 1, Ho = H - 1, Wo = W - 1;
  const size_t n = static_cast<size_t>(B) * Do * Ho * Wo;
  const int C = nd;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = idx % Wo, y = (idx / Wo) % Ho;
    const int z = (idx / (static_cast<size_t>(Wo) * Ho)) % Do;
    const int b = idx / (static_cast<size_t>(Wo) * Ho * Do);
    const size_t p0 = ((static_cast<size_t>(b) * D + z) * H + y) * W + x;
    float r = (ldf(v + (p0 + 1) * C) - ldf(v + p0 * C)) + (ldf(v + (p0 + W) * C + 1) - ldf(v + p0 * C + 1));
    if (nd == 3) r += ldf(v + (p0 + static_cast<size_t>(H) * W) * C + 2) - ldf(v + p0 * C + 2);
    stf(out + idx, r);
  }
}

// =============================================================================================
// host-side launchers (called by the C-ABI in dfl_api.cu)
// =============================================================================================
static int stencil_grid(const StencilParams& p, int nd) {
  return (nd == 3) ? p.B * p.nseg * p.tiles_y * p.tiles_x : p.B * p.tiles_y * p.tiles_x;
}

static void stencil_plan(int nd, const int64_t* dims, StencilParams& p) {
  if (nd == 3) {
    p.B = dims[0]; p.D = dims[1]; p.H = dims[2]; p.W = dims[3];
    p.tiles_x = (p.W + S3_OX - 1) / S3_OX;
    p.tiles_y = (p.H + S3_OY - 1) / S3_OY;
    // z segments: enough blocks for ~3 waves of 148 SMs x 2 blocks, but >= 16 planes per segment
    int cols = p.B * p.tiles_x * p.tiles_y;
    int want = (148 * 2 * 2 + cols - 1) / cols;
    int nseg = want < 1 ? 1 : want;
    int zseg = (p.D + nseg - 1) / nseg;
    if (zseg < 32) zseg = p.D < 32 ? p.D : 32;
    p.zseg = zseg;
    p.nseg = (p.D + zseg - 1) / zseg;
  } else {
    p.B = dims[0]; p.D = 1; p.H = dims[1]; p.W = dims[2];
    p.tiles_x = (p.W + S2_OX - 1) / S2_OX;
    p.tiles_y = (p.H + S2_OY - 1) / S2_OY;
    p.zseg = 1; p.nseg = 1;
  }
}

// dfl_stencil3_lean.cu: all-fp32 3D fast path (persistent CTAs, <= stencil3d_lean_max_blocks() partial sums)
int stencil3d_lean_max_blocks();
int stencil3d_lean_launch(const float* A, const float* X, float* dA, float* vel, double* partials, int B, int D, int H,
                          int W, float c1, float c2, cudaStream_t st, int* nblk);

size_t stencil_loss_workspace_bytes(int nd, const int64_t* dims) {
  StencilParams p{};
  stencil_plan(nd, dims, p);
  const int blocks = std::max(stencil_grid(p, nd), stencil3d_lean_max_blocks());
  return static_cast<size_t>(blocks) * 2 * sizeof(double);
}

template <typename TP, typename TX_, typename TO>
static int stencil_launch_typed(int nd, const void* pot, const void* x, void* dpot, void* vel, float* loss3,
                                void* workspace, const StencilParams& p, double n1, double n2, float w1, float w2,
                                cudaStream_t st) {
  int grid = stencil_grid(p, nd);
  double* part = static_cast<double*>(workspace);
  int lean_blocks = 0;
  if (nd == 3 && std::is_same<TP, float>::value && std::is_same<TX_, float>::value && std::is_same<TO, float>::value &&
      !getenv("DFL_STENCIL_GENERIC")) {
    const int rc = stencil3d_lean_launch(static_cast<const float*>(pot), static_cast<const float*>(x),
                                         static_cast<float*>(dpot), static_cast<float*>(vel), part, p.B, p.D, p.H, p.W,
                                         p.c1, p.c2, st, &lean_blocks);
    if (rc != DFL_OK) return rc;
  }
  if (lean_blocks > 0) {
    grid = lean_blocks;
  } else if (nd == 3) {
    dim3 blk(S3_TX, S3_TY);
    if (vel)
      stencil3d_fused_kernel<TP, TX_, TO, true><<<grid, blk, 0, st>>>(
          static_cast<const TP*>(pot), static_cast<const TX_*>(x), static_cast<TO*>(dpot), static_cast<TO*>(vel), part, p);
    else
      stencil3d_fused_kernel<TP, TX_, TO, false><<<grid, blk, 0, st>>>(
          static_cast<const TP*>(pot), static_cast<const TX_*>(x), static_cast<TO*>(dpot), nullptr, part, p);
  } else {
    dim3 blk(S2_TX, S2_TY);
    if (vel)
      stencil2d_fused_kernel<TP, TX_, TO, true><<<grid, blk, 0, st>>>(
          static_cast<const TP*>(pot), static_cast<const TX_*>(x), static_cast<TO*>(dpot), static_cast<TO*>(vel), part, p);
    else
      stencil2d_fused_kernel<TP, TX_, TO, false><<<grid, blk, 0, st>>>(
          static_cast<const TP*>(pot), static_cast<const TX_*>(x), static_cast<TO*>(dpot), nullptr, part, p);
  }
  DFL_LAUNCH_OK("stencil_fused_kernel");
  stencil_finalize_kernel<<<1, 256, 0, st>>>(part, grid, 1.0 / n1, 1.0 / n2, w1, w2, loss3);
  DFL_LAUNCH_OK("stencil_finalize_kernel");
  return DFL_OK;
}

int stencil_loss_fwdbwd(int nd, const int64_t* dims, const void* pot, int pot_channels, const void* x, void* dpot,
                        void* vel, float* loss3, void* workspace, float w1, float w2, float grad_scale, int dt_pot,
                        int dt_x, int dpot_channels, cudaStream_t st) {
  DFL_REQUIRE(nd == 2 || nd == 3, "stencil_loss: ndim must be 2 or 3 (got %d)", nd);
  StencilParams p{};
  stencil_plan(nd, dims, p);
  DFL_REQUIRE(p.H >= 2 && p.W >= 2 && (nd == 2 || p.D >= 2), "stencil_loss: every spatial extent must be >= 2");
  DFL_REQUIRE(pot_channels >= (nd == 3 ? 3 : 1), "stencil_loss: potential needs >= %d channels", nd == 3 ? 3 : 1);
  p.pot_cs = pot_channels;
  p.dpot_cs = (nd == 2) ? (dpot_channels > 0 ? dpot_channels : 1) : 3;
  DFL_REQUIRE(nd == 2 || pot_channels == 3, "stencil_loss (3D): the potential must have exactly 3 channels");
  const double vox = static_cast<double>(p.B) * p.D * p.H * p.W;
  const double n1 = vox * nd, n2 = vox * nd * nd;
  p.c1 = static_cast<float>(static_cast<double>(w1) * grad_scale / n1);
  p.c2 = static_cast<float>(static_cast<double>(w2) * grad_scale / n2);
  if (dt_pot == DT_F32 && dt_x == DT_F32)
    return stencil_launch_typed<float, float, float>(nd, pot, x, dpot, vel, loss3, workspace, p, n1, n2, w1, w2, st);
  if (dt_pot == DT_BF16 && dt_x == DT_BF16)
    return stencil_launch_typed<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(nd, pot, x, dpot, vel, loss3, workspace,
                                                                            p, n1, n2, w1, w2, st);
  if (dt_pot == DT_F32 && dt_x == DT_BF16)
    return stencil_launch_typed<float, __nv_bfloat16, float>(nd, pot, x, dpot, vel, loss3, workspace, p, n1, n2, w1,
                                                            w2, st);
  if (dt_pot == DT_BF16 && dt_x == DT_F32)
    return stencil_launch_typed<__nv_bfloat16, float, __nv_bfloat16>(nd, pot, x, dpot, vel, loss3, workspace, p, n1,
                                                                    n2, w1, w2, st);
  set_last_error("stencil_loss: unsupported dtype combination (%d,%d)", dt_pot, dt_x);
  return DFL_ERR_UNSUPPORTED;
}

template <typename T>
static int fwd_stencils_typed(int op, int nd, const int64_t* dims, const void* in, int in_cs, void* out0, void* out1,
                              cudaStream_t st) {
  const int B = dims[0];
  const int D = nd == 3 ? dims[1] : 1, H = dims[nd - 1], W = dims[nd];
  const size_t n = static_cast<size_t>(B) * D * H * W;
  const int threads = 256;
  const int grid = static_cast<int>(std::min<size_t>((n + threads - 1) / threads, 148 * 16));
  const T* i = static_cast<const T*>(in);
  switch (op) {
    case 0:  // curl
      if (nd == 2)
        curl2d_kernel<T><<<grid, threads, 0, st>>>(i, static_cast<T*>(out0), B, H, W, in_cs);
      else
        jacobian3d_kernel<T><<<grid, threads, 0, st>>>(i, nullptr, static_cast<T*>(out0), B, D, H, W);
      break;
    case 1:  // jacobian (+ vorticity / curl)
      if (nd == 2)
        jacobian2d_kernel<T><<<grid, threads, 0, st>>>(i, static_cast<T*>(out0), static_cast<T*>(out1), B, H, W);
      else
        jacobian3d_kernel<T><<<grid, threads, 0, st>>>(i, static_cast<T*>(out0), static_cast<T*>(out1), B, D, H, W);
      break;
    case 2:  // divergence
      divergence_kernel<T><<<grid, threads, 0, st>>>(i, static_cast<T*>(out0), B, D, H, W, nd);
      break;
    default:
      set_last_error("unknown stencil op %d", op);
      return DFL_ERR_ARG;
  }
  DFL_LAUNCH_OK("fwd_stencil_kernel");
  return DFL_OK;
}

int fwd_stencils(int op, int nd, const int64_t* dims, const void* in, int in_cs, void* out0, void* out1, int dtype,
                 cudaStream_t st) {
  DFL_REQUIRE(nd == 2 || nd == 3, "stencil: ndim must be 2 or 3 (got %d)", nd);
  for (int k = 1; k <= nd; ++k) DFL_REQUIRE(dims[k] >= 2, "stencil: every spatial extent must be >= 2");
  DFL_REQUIRE(!(op == 0 && nd == 3 && in_cs != 3), "curl3d: potential must have exactly 3 channels");
  if (dtype == DT_F32) return fwd_stencils_typed<float>(op, nd, dims, in, in_cs, out0, out1, st);
  if (dtype == DT_BF16) return fwd_stencils_typed<__nv_bfloat16>(op, nd, dims, in, in_cs, out0, out1, st);
  set_last_error("stencil: unsupported dtype %d", dtype);
  return DFL_ERR_UNSUPPORTED;
}

}  // namespace dfl
