// deepfluids_b200 -- weight gradient of the 3x3(x3) 128->128 convolution on tcgen05 tensor cores.
//
// Replaces TF autodiff's Conv2DBackpropFilter / Conv3DBackpropFilterV2 for the layers built by reference
// model.py:26,68 (slim.conv2d / slim.conv3d, ops.py:12-16):
//     dW[tap][ci][co] = sum_p  X[p + tap - 1][ci] * dP[p][co]          (X zero outside the domain = SAME padding)
//
//   GEMM view per tap:  D_tap[M = ci (128), N = co (128)] += A_tap[M, K] * B[N, K]^T,  K = voxels.
//   Both operands are "MN-major": the TMA box of a channels-last tensor lands in smem as K rows (voxels) of
//   128 bytes (64 channels), 128B-swizzled -- exactly the canonical MN-major UMMA layout, so the activation
//   and gradient bricks feed the tensor core with no transpose.  A_tap is the brick of X shifted by the tap
//   offset (TMA zero-fills out-of-bounds = padding); B is the un-shifted brick of dP, shared by the taps.
//   Each CTA owns a group of <= 4 taps (4 x 128 fp32 TMEM columns = all 512) and a slab of voxel bricks
//   (split-K); it accumulates in TMEM across its whole slab and adds its partial dW into the fp32 gradient
//   with red.global.add at the end.
//   Roles: warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2..5 = epilogue.
#include "dfl_common.cuh"
#include <cmath>
#include <cstdlib>

namespace dfl {

constexpr int WG_TILE_K = 128;                        // voxels per brick
constexpr int WG_OP_BYTES = WG_TILE_K * 128 * 2;      // 32 KB: 128 voxels x 128 channels bf16 (two 64-ch boxes)
constexpr int WG_A_SLOTS = 4, WG_B_SLOTS = 2;
constexpr int WG_THREADS = 192;
constexpr int WG_MAX_TAPS = 4;
constexpr int WG_ONES_BYTES = 2048;   // 16 K-rows x 128 B of bf16 1.0: A operand of the bias-gradient MMA
constexpr int WG_SMEM_BYTES = (WG_A_SLOTS + WG_B_SLOTS) * WG_OP_BYTES + WG_ONES_BYTES + 1024 + 256;

struct WgradParams {
  int B, D, H, W;
  int bd, bh, bw;
  int tz, ty, tx, ntiles;
  int kd, kh, kw;
  int taps_per_group, ngroups, nslabs;
  float* dw;          // fp32 gradient block, element (tap, ci, co) at dw[tap*dw_tap_stride + ci*dw_row_stride + co]
  float* db;          // [128] fp32 bias gradient (sum_p dP[p][co]), accumulated into; may be nullptr
  int in_stride;      // 1, or 2 for the weight gradient of a stride-2 convolution (X sampled at 2p + tap - pad)
  int pad;            // tap offset = tap index - pad   (1 for SAME stride 1; pad_before of TF SAME for stride 2)
  int dw_tap_stride, dw_row_stride;   // TF layout [taps][Cin][Cout]: Cin*Cout and Cout
  // brick mode (stride 1, bw >= 8): ONE y-halo'd brick of X per (dz,dx) group and tile; the three dy taps are descriptor
  // windows shifted by whole x-lines (bw rows = a multiple of 1024 B, so every window stays swizzle-atom aligned: an
  // x-halo variant whose windows start on arbitrary rows measured 2x SLOWER MMAs for these MN-major operands)
  int brick;
  int a_half;         // bytes between the two 64-channel halves of a brick (1024-aligned)
  // operand combinations accumulated into the same result (fp32-grade split operands: x_hi*dP_hi + x_lo*dP_hi + x_hi*dP_lo):
  // combination c reads X at batch b + xoff[c] and dP at batch b + poff[c]; bit c of bias_mask = dP of c enters db
  int ncombo, xoff[3], poff[3], bias_mask;
  int vec_red;        // dw rows are 16-byte aligned: reduce with red.global.add.v4.f32
  // deterministic mode (dfl_set_deterministic): instead of red.global.add into dw / db every CTA stores its accumulators to
  // its own WG_PART_FLOATS slot of this workspace; wgrad_reduce_kernel then sums the slabs in a fixed order
  float* partial;
};
constexpr size_t WG_PART_FLOATS = static_cast<size_t>(WG_MAX_TAPS) * 128 * 128 + 128;
constexpr int WG_BRICK_SLOT = 40960;      // >= 2 * bd*(bh+2)*bw*128 for the tile shapes that use brick mode
constexpr int WG_BRICK_SLOTS = 3;

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmP, WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + WG_A_SLOTS * WG_OP_BYTES;
  uint8_t* sOnes = sB + WG_B_SLOTS * WG_OP_BYTES;
  uint8_t* ctrl = sOnes + WG_ONES_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* a_empty = a_full + WG_A_SLOTS;
  uint64_t* b_full = a_empty + WG_A_SLOTS;
  uint64_t* b_empty = b_full + WG_B_SLOTS;
  uint64_t* acc_full = b_empty + WG_B_SLOTS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x % p.ngroups, slab = blockIdx.x / p.ngroups;
  const int ntaps = p.kd * p.kh * p.kw;
  const int tap0 = group * p.taps_per_group;
  const int gtaps = min(p.taps_per_group, ntaps - tap0);
  // this CTA's slab of bricks: [t_begin, t_end)
  const int nvt = p.ntiles * p.ncombo;          // virtual bricks = (brick, operand combination), combination fastest
  const int per = (nvt + p.nslabs - 1) / p.nslabs;
  const int t_begin = slab * per, t_end = min(nvt, t_begin + per);
  // A tap group with a free 128-column TMEM slot also accumulates the bias gradient db[co] = sum_p dP[p][co] as one more
  // GEMM, ones[M x K] (x) dP -- every row of that accumulator equals db.  Brick mode: every group has 3 taps, so the
  // bricks' bias MMAs are dealt round-robin over the groups (brick % ngroups == group).  Giving all of them to one group
  // made that group's CTAs issue 32 instead of 24 MMAs per brick: they finished last (kernel time = slowest CTA), fell
  // hundreds of bricks behind their slab's other groups and re-read both operands from HBM instead of L2.
  // Tap-list mode (4 taps per group): only the last group (27 = 6*4+3, 9 = 2*4+1) has the free slot.
  const bool bias_rr = p.brick != 0;
  const bool do_bias = (p.db != nullptr) && (gtaps < WG_MAX_TAPS) && (bias_rr || group == p.ngroups - 1);
  for (int i = threadIdx.x; i < WG_ONES_BYTES / 4; i += WG_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmP);
    for (int s = 0; s < WG_A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < WG_B_SLOTS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();     // the ones tile was written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (t_begin < t_end) {
    if (warp == 0) {
      if (lane == 0) {
        uint32_t ia = 0, ib = 0;
        for (int vt = t_begin; vt < t_end; ++vt, ++ib) {
          const int tile = vt / p.ncombo, combo = vt - tile * p.ncombo;
          // z-fastest traversal: a CTA's consecutive bricks are z-neighbours, so the planes the dz = 0/1/2 tap groups
          // share are re-read from L2 within a few bricks instead of one z-plane (4 MB per slab at 128^3) later --
          // ncu showed 13.0 GB of DRAM reads per launch (3x the algorithmic 4.3 GB) with the x-fastest order
          int r = tile;
          const int z0 = (r % p.tz) * p.bd; r /= p.tz;
          const int x0 = (r % p.tx) * p.bw; r /= p.tx;
          const int y0 = (r % p.ty) * p.bh; r /= p.ty;
          const int b = r + p.xoff[combo], bp = r + p.poff[combo];
          {  // B = dP brick (two 64-channel boxes)
            const uint32_t s = ib % WG_B_SLOTS, ph = (ib / WG_B_SLOTS) & 1;
            mbar_wait(&b_empty[s], ph ^ 1);
            mbar_expect_tx(&b_full[s], WG_OP_BYTES);
            tma_load_5d(sB + s * WG_OP_BYTES, &tmP, &b_full[s], 0, x0, y0, z0, bp);
            tma_load_5d(sB + s * WG_OP_BYTES + WG_OP_BYTES / 2, &tmP, &b_full[s], 64, x0, y0, z0, bp);
          }
          if (p.brick) {
            const int dx = group % 3, dz = group / 3;
            const uint32_t s = ia % WG_BRICK_SLOTS, ph = (ia / WG_BRICK_SLOTS) & 1;
            mbar_wait(&a_empty[s], ph ^ 1);
            mbar_expect_tx(&a_full[s], 2 * p.bd * (p.bh + 2) * p.bw * 128);
            const int xs = x0 + dx - p.pad, zs = (p.kd > 1) ? z0 + dz - p.pad : 0;
            tma_load_5d(sA + s * WG_BRICK_SLOT, &tmX, &a_full[s], 0, xs, y0 - p.pad, zs, b);
            tma_load_5d(sA + s * WG_BRICK_SLOT + p.a_half, &tmX, &a_full[s], 64, xs, y0 - p.pad, zs, b);
            ++ia;
          } else
          for (int t = 0; t < gtaps; ++t, ++ia) {
            const int tap = tap0 + t;
            const int dx = tap % p.kw, dy = (tap / p.kw) % p.kh, dz = tap / (p.kw * p.kh);
            const uint32_t s = ia % WG_A_SLOTS, ph = (ia / WG_A_SLOTS) & 1;
            mbar_wait(&a_empty[s], ph ^ 1);
            mbar_expect_tx(&a_full[s], WG_OP_BYTES);
            const int xs = x0 * p.in_stride + dx - p.pad, ys = y0 * p.in_stride + dy - p.pad,
                      zs = (p.kd > 1) ? z0 * p.in_stride + dz - p.pad : 0;
            tma_load_5d(sA + s * WG_OP_BYTES, &tmX, &a_full[s], 0, xs, ys, zs, b);
            tma_load_5d(sA + s * WG_OP_BYTES + WG_OP_BYTES / 2, &tmX, &a_full[s], 64, xs, ys, zs, b);
          }
        }
      }
    } else if (warp == 1) {
      const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem_base);
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 1, 1);   // both operands MN-major
        uint32_t ia = 0, ib = 0;
        // brick mode: byte offset of K step k inside the brick for dy = 0 (tile-independent; computed ONCE -- this thread
        // issues every MMA, so per-MMA integer divisions would cap the issue rate below the tensor pipe's)
        uint32_t koff[WG_TILE_K / 16];
#pragma unroll
        for (int k = 0; k < WG_TILE_K / 16; ++k) {
          const int v = k * 16, l = v / p.bw, xoff = v % p.bw;
          koff[k] = static_cast<uint32_t>((((l / p.bh) * (p.bh + 2) + (l % p.bh)) * p.bw + xoff) * 128) >> 4;   // 16-byte units
        }
        const uint32_t line16 = static_cast<uint32_t>(p.bw * 128) >> 4;
        // The D address uses a uniform-register copy of the TMEM base and the descriptors are advanced by adding to their
        // 16-byte address field: with `tmem_base` (a shared-memory load) in the address the compiler wrapped every
        // tcgen05.mma in an ELECT / R2UR.BROADCAST loop, and rebuilding both descriptors per MMA cost ~10 more
        // uniform-pipe instructions -- this single thread's instruction stream, not the tensor pipe, was the limiter
        // (see dfl_conv_tc.cu).
        const uint32_t tmem0 = tmem_u;
        int combo = t_begin % p.ncombo;
        int bphase = bias_rr ? (t_begin / p.ncombo) % p.ngroups : group;     // brick index mod ngroups, kept incrementally
        bool bias_first = true;
        for (int tile = t_begin; tile < t_end; ++tile, ++ib) {      // tile = virtual brick index here
          const uint32_t sbs = ib % WG_B_SLOTS, bph = (ib / WG_B_SLOTS) & 1;
          mbar_wait(&b_full[sbs], bph);
          tc_fence_after();
          const uint32_t sb = smem_u32(sB + sbs * WG_OP_BYTES);
          if (p.brick) {
            const uint32_t s = ia % WG_BRICK_SLOTS, ph = (ia / WG_BRICK_SLOTS) & 1;
            mbar_wait(&a_full[s], ph);
            tc_fence_after();
            const uint64_t da0 = umma_desc_sw128(smem_u32(sA + s * WG_BRICK_SLOT), p.a_half, 1024);
            const uint64_t db0 = umma_desc_sw128(sb, WG_OP_BYTES / 2, 1024);
            const uint32_t acc = (tile != t_begin) ? 1u : 0u;
#pragma unroll
            for (int t = 0; t < 3; ++t) {                                 // t = dy: window shifted by t x-lines
#pragma unroll
              for (int k = 0; k < WG_TILE_K / 16; ++k)
                umma_bf16(tmem0 + t * 128, da0 + (koff[k] + t * line16), db0 + k * 128, idesc, (k != 0) ? 1u : acc);
            }
            umma_commit(&a_empty[s]);
            ++ia;
          } else
          for (int t = 0; t < gtaps; ++t, ++ia) {
            const uint32_t s = ia % WG_A_SLOTS, ph = (ia / WG_A_SLOTS) & 1;
            mbar_wait(&a_full[s], ph);
            tc_fence_after();
            const uint64_t da0 = umma_desc_sw128(smem_u32(sA + s * WG_OP_BYTES), WG_OP_BYTES / 2, 1024);
            const uint64_t db0 = umma_desc_sw128(sb, WG_OP_BYTES / 2, 1024);
            const uint32_t d_tmem = tmem0 + t * 128;
#pragma unroll
            for (int k = 0; k < WG_TILE_K / 16; ++k) {
              // 16 voxels (K) = 2 groups of 8 rows x 128 B (+2048 B = +128 address units); the 64-channel halves are
              // WG_OP_BYTES/2 apart
              umma_bf16(d_tmem, da0 + k * 128, db0 + k * 128, idesc, (tile != t_begin || k != 0) ? 1u : 0u);
            }
            umma_commit(&a_empty[s]);
          }
          if (do_bias && bphase == group && ((p.bias_mask >> combo) & 1)) {
            const uint32_t so = smem_u32(sOnes);
            const uint64_t dones = umma_desc_sw128(so, 0, 1024), dbb = umma_desc_sw128(sb, WG_OP_BYTES / 2, 1024);
#pragma unroll
            for (int k = 0; k < WG_TILE_K / 16; ++k)
              umma_bf16(tmem0 + 3 * 128, dones, dbb + k * 128, idesc, (!bias_first || k != 0) ? 1u : 0u);
            bias_first = false;
          }
          if (++combo == p.ncombo) {
            combo = 0;
            if (bias_rr && ++bphase == p.ngroups) bphase = 0;
          }
          umma_commit(&b_empty[sbs]);
        }
        umma_commit(acc_full);
      }
    } else {
      const int quarter = warp & 3;
      const int ci = quarter * 32 + lane;
      bool bias_any = false;                 // did any of this CTA's virtual bricks feed the bias accumulator?
      if (do_bias && quarter == 0) {         // same walk as the MMA issuer (these warps idle until acc_full anyway)
        int combo = t_begin % p.ncombo, bphase = bias_rr ? (t_begin / p.ncombo) % p.ngroups : group;
        for (int vt = t_begin; vt < t_end && !bias_any; ++vt) {
          bias_any = (bphase == group) && (((p.bias_mask >> combo) & 1) != 0);
          if (++combo == p.ncombo) {
            combo = 0;
            if (bias_rr && ++bphase == p.ngroups) bphase = 0;
          }
        }
      }
      mbar_wait(acc_full, 0);
      tc_fence_after();
      float* part = p.partial ? p.partial + static_cast<size_t>(blockIdx.x) * WG_PART_FLOATS : nullptr;
      for (int t = 0; t < gtaps; ++t) {
        const int tap = p.brick ? ((group / 3) * 3 + t) * 3 + (group % 3) : tap0 + t;   // brick: group = (dz,dx), t = dy
        float* dst = part ? part + (static_cast<size_t>(t) * 128 + ci) * 128
                          : p.dw + static_cast<size_t>(tap) * p.dw_tap_stride + static_cast<size_t>(ci) * p.dw_row_stride;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t rr[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 128 + c0, rr);
          tmem_ld_wait();
          if (part) {
#pragma unroll
            for (int k = 0; k < 32; k += 4)
              *reinterpret_cast<uint4*>(dst + c0 + k) = make_uint4(rr[k], rr[k + 1], rr[k + 2], rr[k + 3]);
          } else if (p.vec_red) {
#pragma unroll
            for (int k = 0; k < 32; k += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + c0 + k), "f"(__uint_as_float(rr[k])),
                           "f"(__uint_as_float(rr[k + 1])), "f"(__uint_as_float(rr[k + 2])), "f"(__uint_as_float(rr[k + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) atomicAdd(dst + c0 + k, __uint_as_float(rr[k]));
          }
        }
      }
      if (part && p.db && quarter == 0 && !(do_bias && bias_any)) {      // deterministic mode: every CTA publishes a bias row
        for (int c = lane; c < 128; c += 32) part[WG_PART_FLOATS - 128 + c] = 0.f;
      }
      if (do_bias && quarter == 0 && bias_any) {      // row 0 of the bias accumulator (all rows are equal)
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t rr[32];
          tmem_ld_32x32(tmem_base + 3 * 128 + c0, rr);
          tmem_ld_wait();
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              if (part) part[WG_PART_FLOATS - 128 + c0 + k] = __uint_as_float(rr[k]);
              else atomicAdd(p.db + c0 + k, __uint_as_float(rr[k]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Deterministic mode, second pass: dw[tap][ci][:] += sum over the non-empty slabs, in slab order, of the owning tap group's
// partial accumulators; the last block does the same for db over all CTAs.  One block per (tap, ci) row, thread = co.
__global__ void __launch_bounds__(128) wgrad_reduce_kernel(WgradParams p, int nslabs_eff) {
  const int ntaps = p.kd * p.kh * p.kw, co = threadIdx.x;
  const int row = blockIdx.x;
  if (row < ntaps * 128) {
    const int tap = row >> 7, ci = row & 127;
    int group, t;
    if (p.brick) {                       // tap = (dz*3 + dy)*3 + dx, group = dz*3 + dx, t = dy
      const int dx = tap % 3, dy = (tap / 3) % 3, dz = tap / 9;
      group = dz * 3 + dx; t = dy;
    } else {
      group = tap / p.taps_per_group; t = tap - group * p.taps_per_group;
    }
    float acc = 0.f;
    for (int s = 0; s < nslabs_eff; ++s)
      acc += p.partial[static_cast<size_t>(s * p.ngroups + group) * WG_PART_FLOATS + (static_cast<size_t>(t) * 128 + ci) * 128 + co];
    p.dw[static_cast<size_t>(tap) * p.dw_tap_stride + static_cast<size_t>(ci) * p.dw_row_stride + co] += acc;
  } else if (p.db) {
    float acc = 0.f;
    for (int c = 0; c < nslabs_eff * p.ngroups; ++c) acc += p.partial[static_cast<size_t>(c) * WG_PART_FLOATS + WG_PART_FLOATS - 128 + co];
    p.db[co] += acc;
  }
}

// process-wide switch (dfl_set_deterministic): a caller-owned workspace of dfl_deterministic_workspace_bytes() bytes
static float* g_det_ws = nullptr;
static size_t g_det_bytes = 0;
size_t deterministic_workspace_bytes() { return static_cast<size_t>(num_sms()) * WG_PART_FLOATS * sizeof(float); }
int set_deterministic(void* ws, size_t bytes) {
  if (ws) DFL_REQUIRE(bytes >= deterministic_workspace_bytes(), "set_deterministic: workspace of %zu bytes, need %zu", bytes,
                      deterministic_workspace_bytes());
  g_det_ws = static_cast<float*>(ws);
  g_det_bytes = ws ? bytes : 0;
  return DFL_OK;
}
float* deterministic_workspace(size_t* bytes) {
  if (bytes) *bytes = g_det_bytes;
  return g_det_ws;
}

static void pick_brick_w(int D, int H, int W, int& bd, int& bh, int& bw) {
  auto p2 = [](int v) { int r = 1; while (r < v) r <<= 1; return r; };
  bw = std::min(p2(W), 16);
  bh = std::min(p2(H), 128 / bw);
  bd = 128 / (bw * bh);
  if (D == 1) {
    while (bd > 1) { if (bw < 128) bw <<= 1; else bh <<= 1; bd >>= 1; }
  } else if (bd > p2(D)) {
    while (bd > p2(D)) { bw <<= 1; bd >>= 1; }
  }
}

// x: [B, xD, xH, xW, 128] bf16 (one 128-channel block), dpre: [B, D, H, W, 128] bf16 (tile domain).  For stride 1
// x dims == dims; for the gradient of a stride-2 conv x is the fine grid and the TMA walks it with element stride 2.
int wgrad_tc_launch(const void* x, const void* dpre, float* dw, float* db, const int64_t* x_dims, const int64_t* dims,
                    int nd, int in_stride, int pad, int dw_tap_stride, int dw_row_stride, cudaStream_t st, int split,
                    int ksize) {
  DFL_REQUIRE(nd == 2 || nd == 3, "wgrad_tc: ndim must be 2 or 3");
  DFL_REQUIRE(in_stride == 1 || in_stride == 2, "wgrad_tc: in_stride must be 1 or 2");
  WgradParams p{};
  p.B = static_cast<int>(dims[0]);
  p.D = nd == 3 ? static_cast<int>(dims[1]) : 1;
  p.H = static_cast<int>(dims[nd - 1]);
  p.W = static_cast<int>(dims[nd]);
  // ksize = 4 (with in_stride 2, pad 1): the 4x4(x4) stride-2 correlation behind the weight gradient of a phase-decomposed
  // upsample-conv (tap offsets -1 .. 2 on the fine grid)
  DFL_REQUIRE(ksize == 3 || ksize == 4, "wgrad_tc: kernel size must be 3 or 4");
  p.kd = nd == 3 ? ksize : 1;
  p.kh = ksize;
  p.kw = ksize;
  pick_brick_w(p.D, p.H, p.W, p.bd, p.bh, p.bw);
  p.tx = (p.W + p.bw - 1) / p.bw;
  p.ty = (p.H + p.bh - 1) / p.bh;
  p.tz = (p.D + p.bd - 1) / p.bd;
  p.ntiles = p.B * p.tz * p.ty * p.tx;
  const int ntaps = p.kd * p.kh * p.kw;
  p.brick = (in_stride == 1 && (p.bw == 8 || p.bw % 16 == 0) && (p.bw >= 16 || p.bh % 2 == 0) &&
             2 * p.bd * (p.bh + 2) * p.bw * 128 <= WG_BRICK_SLOT) ? 1 : 0;
  p.a_half = p.bd * (p.bh + 2) * p.bw * 128;      // rows * 128 B; bw is a multiple of 8 -> 1024-aligned
  p.taps_per_group = p.brick ? 3 : ((nd == 3) ? 4 : 3);
  p.ngroups = (ntaps + p.taps_per_group - 1) / p.taps_per_group;
  // split != 0: x and dpre are (hi, lo) bf16 pairs stacked along the batch axis ([2B, ...]); three operand combinations
  if (split) {
    p.ncombo = 3;
    p.xoff[0] = 0; p.poff[0] = 0;        // x_hi * dP_hi
    p.xoff[1] = p.B; p.poff[1] = 0;      // x_lo * dP_hi
    p.xoff[2] = 0; p.poff[2] = p.B;      // x_hi * dP_lo
    p.bias_mask = 0b101;                 // db = sum(dP_hi) + sum(dP_lo)
  } else {
    p.ncombo = 1;
    p.bias_mask = 1;
  }
  // slabs: every CTA ends with a (taps x 64 KB) fp32 reduction into dw whose cost grows with the number of slabs while
  // the main loop shrinks with it: T ~ a*nvt/ns + b*ns -> ns ~ sqrt(K * nvt).  With the vectorised reduction
  // (red.global.add.v4.f32; the scalar atomics cost a ~55 us floor per launch) the sweep in
  // profiles/r01_wgrad_slabs.txt is flat for K >= 4; K = 8 only trims the smallest levels.
  {
    const int nvt = p.ntiles * p.ncombo;
    const int cap = std::max(1, num_sms() / p.ngroups);
    static const double kslab = getenv("DFL_WGRAD_SLAB_K") ? atof(getenv("DFL_WGRAD_SLAB_K")) : 8.0;
    int ns = static_cast<int>(std::lround(std::sqrt(kslab * nvt)));
    p.nslabs = std::max(1, std::min(std::min(nvt, cap), ns));
  }
  p.vec_red = (dw_row_stride % 4 == 0 && dw_tap_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0) ? 1 : 0;
  p.dw = dw;
  p.db = db;
  p.in_stride = in_stride;
  p.pad = pad;
  p.dw_tap_stride = dw_tap_stride;
  p.dw_row_stride = dw_row_stride;

  CUtensorMap tmX, tmP;
  {
    const int xD = nd == 3 ? static_cast<int>(x_dims[1]) : 1, xH = static_cast<int>(x_dims[nd - 1]),
              xW = static_cast<int>(x_dims[nd]);
    const uint32_t s = static_cast<uint32_t>(in_stride);
    const uint64_t gd[5] = {128, static_cast<uint64_t>(xW), static_cast<uint64_t>(xH), static_cast<uint64_t>(xD),
                            static_cast<uint64_t>(x_dims[0]) * (split ? 2 : 1)};
    const uint64_t gs[4] = {256, 256ull * xW, 256ull * xW * xH, 256ull * xW * xH * xD};
    const uint32_t box[5] = {64, p.bw * s, p.brick ? static_cast<uint32_t>(p.bh + 2) : p.bh * s,
                             (xD == 1 ? 1u : p.bd * s), 1};
    const uint32_t es[5] = {1, s, s, (xD == 1 ? 1u : s), 1};
    int rc = encode_tensor_map(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x, gd, gs, box, CU_TENSOR_MAP_SWIZZLE_128B, es);
    if (rc) return rc;
  }
  {
    const uint64_t gd[5] = {128, static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H), static_cast<uint64_t>(p.D),
                            static_cast<uint64_t>(p.B) * (split ? 2 : 1)};
    const uint64_t gs[4] = {256, 256ull * p.W, 256ull * p.W * p.H, 256ull * p.W * p.H * p.D};
    const uint32_t box[5] = {64, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh), static_cast<uint32_t>(p.bd), 1};
    int rc = encode_tensor_map(&tmP, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dpre, gd, gs, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    DFL_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
    attr_set = true;
  }
  p.partial = g_det_ws;
  DFL_REQUIRE(!p.partial || static_cast<size_t>(p.ngroups) * p.nslabs * WG_PART_FLOATS * sizeof(float) <= g_det_bytes,
              "wgrad_tc: deterministic workspace too small for %d CTAs", p.ngroups * p.nslabs);
  wgrad_tc_kernel<<<p.ngroups * p.nslabs, WG_THREADS, WG_SMEM_BYTES, st>>>(tmX, tmP, p);
  DFL_LAUNCH_OK("wgrad_tc_kernel");
  if (p.partial) {
    const int nvt = p.ntiles * p.ncombo, per = (nvt + p.nslabs - 1) / p.nslabs;
    const int nslabs_eff = (nvt + per - 1) / per;                 // slabs that own at least one brick
    wgrad_reduce_kernel<<<ntaps * 128 + (p.db ? 1 : 0), 128, 0, st>>>(p, nslabs_eff);
    DFL_LAUNCH_OK("wgrad_reduce_kernel");
  }
  return DFL_OK;
}

}  // namespace dfl
