// deepfluids_b200 -- 3x3(x3) convolution, Cin = multiple of 64, Cout = 128, as an implicit GEMM on the
// 5th-gen tensor cores (tcgen05.mma, fp32 accumulators in TMEM, operands staged by TMA).
//
// Replaces slim.conv2d/conv3d (+ bias-add + leaky-ReLU + residual add + nearest x2 upsample, which TF1 runs
// as separate full passes) on the generator/AE path: reference ops.py:12-16 called from model.py:26,68 and
// the residual/upsample glue model.py:34-36,76-79.  The same kernel is the data-gradient pass (dgrad): a
// convolution of dL/dy with the tap-flipped, channel-transposed weights, whose epilogue applies the
// leaky-ReLU derivative of the producing layer and/or adds the residual-branch gradient.
//
//   GEMM view:  D[M = 128 output voxels (a 3D brick), N = 128 out-channels] += A[M, K] * B[N, K]^T
//               K = taps x Cin, walked tap by tap in 64-channel slices (one 128-byte swizzle row each).
//   A operand:  for tap (dz,dy,dx) the TMA loads the brick shifted by (dz-1,dy-1,dx-1) straight out of the
//               NDHWC activation tensor (5-D tiled tensor map, box = [64 ch, bw, bh, bd, 1]); out-of-bounds
//               voxels are zero-filled by the TMA unit, which *is* TF's SAME padding -- no im2col buffer, no
//               padded copy.  The box lands in shared memory as 128 rows x 128 B, 128B-swizzled = the canonical
//               K-major UMMA layout.
//   B operand:  pre-packed bf16 weights [Cout][tap*Cin + ci] (K-major), 2-D tensor map, box [64, 128].
//   Roles:      warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
//               warps 2..5 = epilogue (TMEM -> registers -> bias/lrelu/mask/residual -> bf16 -> global,
//               optionally replicated 2x2(x2) = fused nearest-neighbour upsample).
//   Pipelines:  6-stage smem ring (full/empty mbarriers), 2 TMEM accumulators (full/empty mbarriers) so the
//               epilogue of tile i overlaps the MMAs of tile i+1; persistent CTAs, static tile schedule.
//
// v2 ("tap windows", default): measured on B200 (tools/umma_probe.cu, profiles/r01_umma_probe.txt) the tensor core
// applies the 128B swizzle to ABSOLUTE shared-memory address bits, so a UMMA descriptor may start at any 128-byte row
// of a TMA-written image and use any row pitch.  The kernel therefore keeps ONE halo'd activation brick resident in
// smem per 64-channel slice and reads every (dz,dy,dx) tap as a shifted descriptor window over it (no per-tap
// reload: A traffic / 9..27), processes 256 voxels per tile as two M=128 halves that share each streamed weight
// tile (B traffic / 2), and spends the freed smem on a 6..8-deep weight ring (hides the ~2500-cycle TMA latency that
// bound v1: profiles/r01_ncu_conv_tc_v1_c2_fullres.json).
#include <stdlib.h>

#include "dfl_common.cuh"
#ifndef DFL_MMA_ISSUE
#define DFL_MMA_ISSUE 0
#endif
#ifndef DFL_DESC_INC
#define DFL_DESC_INC 1
#endif

namespace dfl {

constexpr int CT_BLOCK_M = 128;
constexpr int CT_BLOCK_N = 128;
constexpr int CT_BLOCK_K = 64;                    // bf16 elements = 128 bytes = one swizzle row
constexpr int CT_STAGES = 6;
constexpr int CT_A_BYTES = CT_BLOCK_M * CT_BLOCK_K * 2;   // 16 KB
constexpr int CT_B_BYTES = CT_BLOCK_N * CT_BLOCK_K * 2;   // 16 KB
constexpr int CT_STAGE_BYTES = CT_A_BYTES + CT_B_BYTES;
constexpr int CT_THREADS = 192;
constexpr int CT_SMEM_BYTES = CT_STAGES * CT_STAGE_BYTES + 1024 /*align slack*/ + 1024 /*barriers, bias*/;
constexpr int CT_STAGES_PAIR = 4;                 // paired bricks: 4 x (2 x 16 KB activations + 16 KB weights)
constexpr int CT_SMEM_BYTES_PAIR = CT_STAGES_PAIR * (2 * CT_A_BYTES + CT_B_BYTES) + 1024 + 1024;

enum : int { CF_LRELU = 1, CF_OUT2_UPSAMPLE = 2, CF_MASK_AFTER_RESIDUAL = 4, CF_SPLIT_IO = 8 };
constexpr int CT_MAX_TAPS = 64;                    // 27 for a 3x3x3 layer; 64 = 8 output phases x 2x2x2 taps of the
                                                  // data gradient of a phase-decomposed upsample-conv (engine.py)

struct ConvTcParams {
  int B, D, H, W;
  int bd, bh, bw;             // brick, bd*bh*bw == 128
  int tz, ty, tx, ntiles;     // tiles per axis / total (over batch too)
  int kd, kh, kw;             // 3,3,3 (3D) or 1,3,3 (2D)
  int cin_chunks;             // Cin / 64
  int flags;
  const float* bias;          // [128] or nullptr
  const __nv_bfloat16* mask_src;  // multiply by lrelu'(mask_src) (dgrad) or nullptr
  const __nv_bfloat16* residual;  // added into out2 or nullptr
  __nv_bfloat16* out;         // [B,D,H,W,128] or nullptr
  __nv_bfloat16* out2;        // [B,D,H,W,128] or upsampled [B,(2D),2H,2W,128] or nullptr
  float* out_f32;             // kN = 16 variant: fp32 [B,D,H,W,cout_small]
  int cout_small;
  // output tensor geometry (== B,D,H,W of the tile domain unless the generic tap kernel maps positions)
  int oD, oH, oW;
  // virtual -> physical channel block of the input (64-channel slice c reads block blkmap[c >> 1]).  Identity except in
  // the fp32-grade "bf16x3" mode, where the input is a (hi, lo) pair and the virtual blocks are [hi, lo, hi].
  int blkmap[8];
  // ---- generic per-tap kernel only (strided / transposed-strided convolutions, explicit tap lists) ----
  int in_stride;              // A box origin = in_stride * tile origin + tap offset
  int out_stride, orz, ory, orx;   // output voxel = out_stride * tile voxel + (orz, ory, orx)
  int ntap;
  int tap_dz[CT_MAX_TAPS], tap_dy[CT_MAX_TAPS], tap_dx[CT_MAX_TAPS];   // input-coordinate offsets
  int tap_col[CT_MAX_TAPS];   // weight K-column (elements) of the tap's first 64-channel slice
};

__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = __uint_as_float(w[k] << 16);
    f[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// The shared epilogue (defined below) converts one accumulator row to bf16 outputs.  EpiPre carries the global-memory
// operands of one 32-channel chunk (lrelu-mask source, residual) requested one chunk ahead of their use.
struct EpiPre { uint4 m[4], r[4]; };
__device__ __forceinline__ size_t epi_pos(const ConvTcParams& p, int b, int z, int y, int x) {
  return ((static_cast<size_t>(b) * p.oD + z) * p.oH + y) * p.oW + x;
}
__device__ __forceinline__ void epi_prefetch(const ConvTcParams& p, bool valid, size_t pos, int c0, EpiPre& e);
__device__ __forceinline__ void conv_epilogue_row(const ConvTcParams& p, uint32_t taddr, bool valid, int b, int z,
                                                  int y, int x, const float* s_bias, EpiPre& pre, bool next_valid,
                                                  size_t next_pos);

// =============================================================================================
// Generic per-tap kernel (one TMA box per (tap, 64-channel slice); 6 x 32 KB stages).  Used where the tap-window
// kernel does not apply: stride-2 convolutions (TMA element strides pick every 2nd voxel), their data gradient
// (one launch per output parity class with that class's tap subset, outputs written to the strided positions) and
// any explicit tap list.  Inputs wider than 128 channels are stored as channel blocks [nblk*B, D, H, W, 128]:
// 64-channel slice c lives in block c>>1, i.e. at batch coordinate b + (c>>1)*B.
// =============================================================================================
// kPair: the CTA processes TWO bricks (tiles 2q, 2q+1) per step of its schedule; they share every streamed weight tile, so
// a stage carries 2 x 16 KB of activations + 16 KB of weights for 8 MMAs instead of 16 + 16 KB for 4: 96 instead of 128
// bytes of L2->SM traffic per tensor-core clock.  ncu on the phase-decomposed launches of the 128^3 step showed this
// kernel pulling 58-67 B/clk/SM from L2 with the tensor pipe 46-53 % active -- operand delivery, not the epilogue, is its
// limit (profiles/r02_ncu_conv_tap.txt).  Used when the launch has at least two bricks per SM.
template <bool kPair>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_tap_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvTcParams p) {
  constexpr int NA = kPair ? 2 : 1;
  constexpr int STAGES = kPair ? CT_STAGES_PAIR : CT_STAGES;
  constexpr int STAGE_BYTES = NA * CT_A_BYTES + CT_B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ctrl = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);            // [STAGES]
  uint64_t* empty_bar = full_bar + CT_STAGES;                        // [STAGES]
  uint64_t* tfull_bar = empty_bar + CT_STAGES;                       // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                              // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);  // [1]
  float* s_bias = reinterpret_cast<float*>(ctrl + 256);              // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = p.ntap * p.cin_chunks;
  const int nsteps = kPair ? (p.ntiles + 1) / 2 : p.ntiles;          // schedule steps (brick pairs / bricks)

  if (threadIdx.x < CT_BLOCK_N) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 256 * NA);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int step = blockIdx.x; step < nsteps; step += gridDim.x) {
        int x0[NA], y0[NA], z0[NA], bb[NA];
#pragma unroll
        for (int h = 0; h < NA; ++h) {
          int r = min(NA * step + h, p.ntiles - 1);       // an odd brick count: the ghost half re-loads the last brick
          x0[h] = (r % p.tx) * p.bw; r /= p.tx;
          y0[h] = (r % p.ty) * p.bh; r /= p.ty;
          z0[h] = (r % p.tz) * p.bd; r /= p.tz;
          bb[h] = r;
        }
        for (int t = 0; t < p.ntap; ++t)
          for (int c = 0; c < p.cin_chunks; ++c, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
            uint8_t* sa = smem + s * STAGE_BYTES;
#pragma unroll
            for (int h = 0; h < NA; ++h)
              tma_load_5d(sa + h * CT_A_BYTES, &tmA, &full_bar[s], (c & 1) * CT_BLOCK_K, x0[h] * p.in_stride + p.tap_dx[t],
                          y0[h] * p.in_stride + p.tap_dy[t], z0[h] * p.in_stride + p.tap_dz[t], bb[h] + p.blkmap[c >> 1] * p.B);
            tma_load_2d(sa + NA * CT_A_BYTES, &tmB, &full_bar[s], p.tap_col[t] + c * CT_BLOCK_K, 0);
          }
      }
    }
  } else if (warp == 1) {
    const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem_base);   // uniform-register copy (see conv_tc2_kernel)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(CT_BLOCK_M, CT_BLOCK_N, 0, 0);
      uint32_t it = 0, tcount = 0;
      for (int step = blockIdx.x; step < nsteps; step += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(&tempty_bar[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + acc * (NA * CT_BLOCK_N);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sb = sa + NA * CT_A_BYTES;
          const uint64_t db0 = umma_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int h = 0; h < NA; ++h) {
            const uint64_t da0 = umma_desc_sw128(sa + h * CT_A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < CT_BLOCK_K / 16; ++k)      // +32 B per K step = +2 address units
              umma_bf16(d_tmem + h * CT_BLOCK_N, da0 + 2 * k, db0 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int lw = row % p.bw, lh = (row / p.bw) % p.bh, ld = row / (p.bw * p.bh);
    uint32_t tcount = 0;
    for (int step = blockIdx.x; step < nsteps; step += gridDim.x, ++tcount) {
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
#pragma unroll 1
      for (int h = 0; h < NA; ++h) {
        const int tile = NA * step + h;
        int r = min(tile, p.ntiles - 1);
        const int x = (r % p.tx) * p.bw + lw; r /= p.tx;
        const int y = (r % p.ty) * p.bh + lh; r /= p.ty;
        const int z = (r % p.tz) * p.bd + ld; r /= p.tz;
        const int b = r;
        const int xo = x * p.out_stride + p.orx, yo = y * p.out_stride + p.ory, zo = z * p.out_stride + p.orz;
        const bool valid = (tile < p.ntiles) && (x < p.W) && (y < p.H) && (z < p.D) && (xo < p.oW) && (yo < p.oH) && (zo < p.oD);
        EpiPre pre;
        epi_prefetch(p, valid, epi_pos(p, b, zo, yo, xo), 0, pre);      // in flight while the tile's MMAs finish
        if (h == 0) {
          mbar_wait(&tfull_bar[acc], aph);
          tc_fence_after();
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + (acc * NA + h) * CT_BLOCK_N;
        conv_epilogue_row(p, taddr, valid, b, zo, yo, xo, s_bias, pre, false, 0);
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 256 * NA);
  }
}

// =============================================================================================
// v2 kernel: resident brick slots + tap windows
//   3D tile = 2(z) x 16(y) x 8(x) voxels; M-half h = z-plane h; brick = 4 plane slots of 18 x 10 rows (128 B each)
//   2D tile = 16(y) x 16(x) voxels;        M-half h = x-half h;  brick = 18 x 18 rows, 2 slots ping-pong per phase
//   phase = (tile, 64-channel slice);  warps: 0 = brick producer, 1 = MMA issuer (+TMEM alloc), 2 = weight producer,
//   3..6 = epilogue (2D, kN = 128: 3..10, two warps per TMEM lane quarter).
// =============================================================================================
constexpr int C2_BSTAGES_MAX = 8;
// kN = 128: one 4 KB transposition image per epilogue warp -- 3D: 4 warps, 7 weight stages; 2D: 8 warps (two per TMEM
// quarter: a 2D tile has a third of the MMA time, and four warps' dependent tmem -> smem -> global chains did not fit
// under it), 6 weight stages.  The N = 16 variant has no image and 8 stages.
#ifndef DFL_EPI_WARPS_3D
#define DFL_EPI_WARPS_3D 4
#endif
__host__ __device__ constexpr int c2_epi_warps(bool k3D, int kN) { return kN != 128 ? 4 : (k3D ? DFL_EPI_WARPS_3D : 8); }
__host__ __device__ constexpr int c2_bstages(bool k3D, int kN) { return kN != 128 ? 8 : (c2_epi_warps(k3D, kN) == 4 ? 7 : 6); }
__host__ __device__ constexpr int c2_threads(bool k3D, int kN) { return 96 + 32 * c2_epi_warps(k3D, kN); }
constexpr int C2_SLOT_BYTES_3D = 23552;   // 180 rows * 128 B = 23040, padded to a 1024-byte multiple
constexpr int C2_SLOT_BYTES_2D = 41984;   // 324 rows * 128 B = 41472, padded
constexpr int C2_BRICK_BYTES = 4 * C2_SLOT_BYTES_3D;   // 94208 >= 2 * C2_SLOT_BYTES_2D (83968)
constexpr int C2_SMEM_BYTES = C2_BRICK_BYTES + 8 * CT_B_BYTES + 1024 + 1024;   // 8 stages, or 7 (6) stages + 16 (32) KB of images

// ---- epilogue helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ void epi_load32(const __nv_bfloat16* ptr, float (&f)[32]) {
  const uint4* m = reinterpret_cast<const uint4*>(ptr);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float t[8];
    unpack_bf16x8(__ldg(m + q), t);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[q * 8 + k] = t[k];
  }
}
__device__ __forceinline__ void epi_unpack32(const uint4 (&q4)[4], float (&f)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float t[8];
    unpack_bf16x8(q4[q], t);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[q * 8 + k] = t[k];
  }
}
__device__ __forceinline__ void epi_prefetch(const ConvTcParams& p, bool valid, size_t pos, int c0, EpiPre& e) {
  if (!valid) return;
  if (p.mask_src) {
    const uint4* m = reinterpret_cast<const uint4*>(p.mask_src + pos * CT_BLOCK_N + c0);
#pragma unroll
    for (int q = 0; q < 4; ++q) e.m[q] = __ldg(m + q);
  }
  if (p.out2 && p.residual) {
    const uint4* r = reinterpret_cast<const uint4*>(p.residual + pos * CT_BLOCK_N + c0);
#pragma unroll
    for (int q = 0; q < 4; ++q) e.r[q] = __ldg(r + q);
  }
}
// store 32 channels as bf16; in split mode also the residual lo = bf16(v - float(bf16(v))) one block further
__device__ __forceinline__ void epi_store32(__nv_bfloat16* ptr, const float (&v)[32], bool split, size_t blkstride) {
  uint4* o = reinterpret_cast<uint4*>(ptr);
  uint32_t w[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) w[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
#pragma unroll
  for (int q = 0; q < 4; ++q) o[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  if (split) {
    uint4* l = reinterpret_cast<uint4*>(ptr + blkstride);
    uint32_t wl[16];
#pragma unroll
    for (int k = 0; k < 16; ++k)
      wl[k] = pack_bf16x2(v[2 * k] - __uint_as_float(w[k] << 16), v[2 * k + 1] - __uint_as_float(w[k] & 0xFFFF0000u));
#pragma unroll
    for (int q = 0; q < 4; ++q) l[q] = make_uint4(wl[4 * q], wl[4 * q + 1], wl[4 * q + 2], wl[4 * q + 3]);
  }
}

// 8 channels (one 16-byte piece) as bf16; split mode: lo = bf16(v - float(hi)) one block further
__device__ __forceinline__ void epi_store8(__nv_bfloat16* ptr, const float (&v)[8], bool split, size_t blkstride) {
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
  *reinterpret_cast<uint4*>(ptr) = make_uint4(w[0], w[1], w[2], w[3]);
  if (split) {
    uint32_t wl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      wl[k] = pack_bf16x2(v[2 * k] - __uint_as_float(w[k] << 16), v[2 * k + 1] - __uint_as_float(w[k] & 0xFFFF0000u));
    *reinterpret_cast<uint4*>(ptr + blkstride) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
  }
}

// One accumulator row (one output voxel, 128 channels) -> outputs:
//   v    = acc + bias;  lrelu if CF_LRELU
//   out  = v * lrelu'(mask_src)                (mask only if given and not CF_MASK_AFTER_RESIDUAL)
//   out2 = (v + residual) [* lrelu'(mask_src) if CF_MASK_AFTER_RESIDUAL], nearest-x2 replicated if CF_OUT2_UPSAMPLE
// CF_SPLIT_IO (fp32-grade mode): every bf16 tensor is a (hi, lo) pair one block (B*voxels*128 elements) apart: residual
// is read as hi + lo, outputs are written as hi = bf16(v), lo = bf16(v - hi); masks use the hi part (same sign).
// The mask / residual operands of chunk c0+32 (or of chunk 0 of the row the caller names as `next`) are requested before
// chunk c0 is processed: with the loads issued on demand every chunk paid a dependent global round trip (8 per 256-voxel
// tile), which exceeded the MMA time of a 2D tile (K = 1152) and left the 2D data-gradient launches epilogue-bound.
__device__ __forceinline__ void conv_epilogue_row(const ConvTcParams& p, uint32_t taddr, bool valid, int b, int z,
                                                  int y, int x, const float* s_bias, EpiPre& pre, bool next_valid,
                                                  size_t next_pos) {
  const bool ups = (p.flags & CF_OUT2_UPSAMPLE) != 0;
  const bool act = (p.flags & CF_LRELU) != 0;
  const bool mask_after = (p.flags & CF_MASK_AFTER_RESIDUAL) != 0;
  const bool split = (p.flags & CF_SPLIT_IO) != 0;
  const size_t vox = static_cast<size_t>(p.oD) * p.oH * p.oW;
  const size_t blk = static_cast<size_t>(p.B) * vox * CT_BLOCK_N;          // block stride of an output-shaped tensor
  const size_t pos = epi_pos(p, b, z, y, x);
#pragma unroll 1
  for (int c0 = 0; c0 < CT_BLOCK_N; c0 += 32) {
    uint32_t rr[32];
    tmem_ld_32x32(taddr + c0, rr);
    const EpiPre cur = pre;
    if (c0 + 32 < CT_BLOCK_N) epi_prefetch(p, valid, pos, c0 + 32, pre);
    else epi_prefetch(p, next_valid, next_pos, 0, pre);
    tmem_ld_wait();
    if (!valid) continue;
    float v[32], m[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      float t = __uint_as_float(rr[k]) + s_bias[c0 + k];
      v[k] = act ? lrelu_f(t) : t;
    }
    if (p.mask_src) {
      epi_unpack32(cur.m, m);
#pragma unroll
      for (int k = 0; k < 32; ++k) m[k] = lrelu_grad_from_out(m[k]);
    }
    if (p.out) {
      float o[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) o[k] = (p.mask_src && !mask_after) ? v[k] * m[k] : v[k];
      epi_store32(p.out + pos * CT_BLOCK_N + c0, o, split, blk);
    }
    if (p.out2) {
      if (p.residual) {
        float f[32];
        epi_unpack32(cur.r, f);
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] += f[k];
        if (split) {
          epi_load32(p.residual + blk + pos * CT_BLOCK_N + c0, f);
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] += f[k];
        }
      }
      if (p.mask_src && mask_after) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] *= m[k];
      }
      if (!ups) {
        epi_store32(p.out2 + pos * CT_BLOCK_N + c0, v, split, blk);
      } else {
        // nearest-neighbour x2 (ops.py:75-91): out[2i+a] = in[i]; the z axis only when the conv is 3D
        const int zr = (p.kd > 1) ? 2 : 1;
        const int D2 = p.oD * zr, H2 = p.oH * 2, W2 = p.oW * 2;
        const size_t blk2 = blk * (zr * 4);
        for (int a = 0; a < zr; ++a)
          for (int e = 0; e < 2; ++e)
            for (int f = 0; f < 2; ++f) {
              const size_t pos2 = ((static_cast<size_t>(b) * D2 + (z * zr + a)) * H2 + (2 * y + e)) * W2 + (2 * x + f);
              epi_store32(p.out2 + pos2 * CT_BLOCK_N + c0, v, split, blk2);
            }
      }
    }
  }
}

// ---- transposed epilogue of the tap-window kernel (kN = 128) ------------------------------------------------
// Same arithmetic as conv_epilogue_row, different thread mapping.  Written straight from the thread that owns a TMEM row,
// every 16-byte global access of a warp touches 32 different 128-byte lines (32 half-filled sectors per instruction); that
// access pattern cost 20 % of the 2D conv time and made the store-heavy launches LSU-bound (store-less timing run:
// profiles/r01_diag_nostore_c2_bf16.json).  Here every epilogue warp transposes its own 32 rows, one 32-channel chunk at a
// time, through a private 4 KB shared-memory image (no CTA-level barrier, only __syncwarp):
//   phase 1 (lane = TMEM row):  v = lrelu(acc + bias) as fp32 -> image [32 rows][128 B], 16-byte chunks XOR-swizzled;
//   phase 2 (lane = 8 channels of a row; 4 lanes = one row's 64-byte chunk = 2 complete sectors, a warp instruction = 8
//            x-adjacent voxels): mask / residual pieces, rounding to bf16 (and the lo part in split mode), stores, x2
//            replication.
// The global operand of phase 2 (the lrelu-mask source, else the residual) is requested one chunk ahead; chunk 0's before
// the wait for the accumulator.  3D: c3 +4.5 %, c4 +2 % over the row-wise epilogue.  In 2D, whose tiles carry a third of a
// 3D tile's MMA time, ncu showed the row-wise epilogue bound by the L1 data pipe (LSU wavefronts 73 % of peak in the data-
// gradient launches: 32 wavefronts per scattered 16-byte request), but with four epilogue warps every transposed variant
// was SLOWER (2D conv per step: row-wise 2.94 ms; this one 3.28 ms; CTA-wide 64-channel images with two named barriers
// per chunk 3.57-3.73 ms): the serial tmem -> smem -> global chain of a chunk, eight times per tile, exceeded the tile's
// MMA time.  The 2D kernel therefore runs EIGHT epilogue warps, two per TMEM quarter, alternating chunks.
struct EpiChunk { uint4 q[4]; };
template <bool k3D>
struct EpiWarpT {
  const ConvTcParams& p;
  int b, z0, ybase, x0, xi, piece;
  bool need_m, need_r;
  __device__ __forceinline__ EpiWarpT(const ConvTcParams& p_, int lane, int b_, int z0_, int ybase_, int x0_)
      : p(p_), b(b_), z0(z0_), ybase(ybase_), x0(x0_), xi(lane >> 2), piece(lane & 3) {
    need_m = p.mask_src != nullptr;
    need_r = p.out2 != nullptr && p.residual != nullptr;
  }
  // chunk i = 4 * half + (32-channel chunk);  phase-2 row of iteration it: line = ybase + it, x = xbase(half) + xi
  __device__ __forceinline__ int cz(int i) const { return k3D ? z0 + (i >> 2) : 0; }
  __device__ __forceinline__ int cx(int i) const { return (k3D ? x0 : x0 + 8 * (i >> 2)) + xi; }
  __device__ __forceinline__ bool ok(int i) const { return cx(i) < p.W && cz(i) < p.D; }
  __device__ __forceinline__ size_t pos0(int i) const {
    return ((static_cast<size_t>(b) * p.oD + cz(i)) * p.oH + ybase) * p.oW + cx(i);
  }
  __device__ __forceinline__ void prefetch(int i, EpiChunk& e) const {
    if (!(need_m || need_r) || !ok(i)) return;
    const __nv_bfloat16* src = need_m ? p.mask_src : p.residual;
    const size_t base = pos0(i) * CT_BLOCK_N + (i & 3) * 32 + piece * 8, st = static_cast<size_t>(p.oW) * CT_BLOCK_N;
#pragma unroll
    for (int it = 0; it < 4; ++it)
      if (ybase + it < p.H) e.q[it] = __ldg(reinterpret_cast<const uint4*>(src + base + it * st));
  }
};

template <bool k3D>
__device__ __forceinline__ void conv_epilogue_warp_t(const ConvTcParams& p, const EpiWarpT<k3D>& T, uint32_t taddr0,
                                                     uint32_t wimg, int lane, const float* s_bias, EpiChunk& pre,
                                                     int first, int step) {
  const bool ups = (p.flags & CF_OUT2_UPSAMPLE) != 0;
  const bool act = (p.flags & CF_LRELU) != 0;
  const bool mask_after = (p.flags & CF_MASK_AFTER_RESIDUAL) != 0;
  const bool split = (p.flags & CF_SPLIT_IO) != 0;
  const size_t vox = static_cast<size_t>(p.oD) * p.oH * p.oW;
  const size_t blk = static_cast<size_t>(p.B) * vox * CT_BLOCK_N;
  const bool need_m = T.need_m, need_r = T.need_r;
#pragma unroll 1
  for (int i = first; i < 8; i += step) {       // this warp's chunks (two warps per TMEM quarter split them in 2D)
    const int c0 = (i & 3) * 32;
    uint32_t rr[32];
    tmem_ld_32x32(taddr0 + (i >> 2) * CT_BLOCK_N + c0, rr);
    const EpiChunk cur = pre;
    if (i + step < 8) T.prefetch(i + step, pre);   // next chunk's operand: in flight across both phases of this one
    tmem_ld_wait();
    // ---- phase 1: this lane's TMEM row, 32 channels
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float t[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = __uint_as_float(rr[q * 4 + e]) + s_bias[c0 + q * 4 + e];
        t[e] = act ? lrelu_f(a) : a;
      }
      const uint32_t o = lane * 128 + ((q ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(wimg + o), "f"(t[0]), "f"(t[1]), "f"(t[2]), "f"(t[3]) : "memory");
    }
    __syncwarp();
    // ---- phase 2
    if (T.ok(i)) {
      const int z = T.cz(i), x = T.cx(i);
      const size_t pos0 = T.pos0(i);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int y = T.ybase + it;
        if (y >= p.H) continue;
        const int r = it * 8 + T.xi;
        float v[8];
        {
          const uint32_t a0 = wimg + r * 128 + (((2 * T.piece) ^ (r & 7)) << 4);
          const uint32_t a1 = wimg + r * 128 + (((2 * T.piece + 1) ^ (r & 7)) << 4);
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a0));
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a1));
        }
        const size_t off = (pos0 + static_cast<size_t>(it) * p.oW) * CT_BLOCK_N + c0 + T.piece * 8;
        float m[8];
        if (need_m) {
          unpack_bf16x8(cur.q[it], m);
#pragma unroll
          for (int k = 0; k < 8; ++k) m[k] = lrelu_grad_from_out(m[k]);
        }
        if (p.out) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = (need_m && !mask_after) ? v[k] * m[k] : v[k];
          epi_store8(p.out + off, o, split, blk);
        }
        if (p.out2) {
          if (need_r) {
            float f[8];
            // the prefetched operand is the residual unless a mask source is present too (AE: mask after residual)
            unpack_bf16x8(need_m ? __ldg(reinterpret_cast<const uint4*>(p.residual + off)) : cur.q[it], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += f[k];
            if (split) {
              unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(p.residual + blk + off)), f);
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] += f[k];
            }
          }
          if (need_m && mask_after) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] *= m[k];
          }
          if (!ups) {
            epi_store8(p.out2 + off, v, split, blk);
          } else {
            // nearest-neighbour x2 (ops.py:75-91): out[2i+a] = in[i]; the z axis only when the conv is 3D
            const int zr = k3D ? 2 : 1;
            const int D2 = p.oD * zr, H2 = p.oH * 2, W2 = p.oW * 2;
            const size_t blk2 = blk * (zr * 4);
#pragma unroll
            for (int a = 0; a < zr; ++a)
#pragma unroll
              for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                  const size_t pos2 = ((static_cast<size_t>(T.b) * D2 + (z * zr + a)) * H2 + (2 * y + e)) * W2 + (2 * x + f);
                  epi_store8(p.out2 + pos2 * CT_BLOCK_N + c0 + T.piece * 8, v, split, blk2);
                }
          }
        }
      }
    }
    __syncwarp();                          // the image may be overwritten
  }
}

// kN = 128: the 128->128 layers.  kN = 16: the 128 -> 1..3 output conv (model.py:42,84), weights zero-padded to 16
// output channels; its epilogue writes fp32 [voxel][p.cout_small] (+ bias) = the network output (potential).
template <bool k3D, int kN>
__global__ void __launch_bounds__(c2_threads(k3D, kN), 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ConvTcParams p) {
  constexpr int NSLOT = k3D ? 4 : 2;
  constexpr int SLOT_BYTES = k3D ? C2_SLOT_BYTES_3D : C2_SLOT_BYTES_2D;
  constexpr int SLOT_TX = (k3D ? 180 : 324) * 128;    // bytes landed per slot fill
  constexpr int HX = k3D ? 10 : 18;                   // brick rows per y-line
  constexpr int TZ = k3D ? 2 : 1, TY = 16, TX = k3D ? 8 : 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem + C2_BRICK_BYTES;
  constexpr int B_BYTES = kN * CT_BLOCK_K * 2;
  constexpr int C2_BSTAGES = c2_bstages(k3D, kN);
  uint8_t* sE = sB + C2_BSTAGES * CT_B_BYTES;             // 3D, kN = 128: the epilogue warps' transposition images
  uint8_t* ctrl = sB + 8 * CT_B_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);   // [4]
  uint64_t* a_empty = a_full + 4;                         // [4]
  uint64_t* b_full = a_empty + 4;                         // [C2_BSTAGES]
  uint64_t* b_empty = b_full + C2_BSTAGES_MAX;            // [C2_BSTAGES]
  uint64_t* tfull_bar = b_empty + C2_BSTAGES_MAX;         // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                   // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(ctrl + 512);   // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntaps = p.kd * p.kh * p.kw;

  if (threadIdx.x < CT_BLOCK_N)
    s_bias[threadIdx.x] = (p.bias && (kN == CT_BLOCK_N || threadIdx.x < p.cout_small)) ? p.bias[threadIdx.x] : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < 4; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C2_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 32 * c2_epi_warps(k3D, kN)); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================================ brick producer ================================
    if (lane == 0) {
      uint32_t fills[NSLOT];
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) fills[s] = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        int r = tile;
        const int x0 = (r % p.tx) * TX; r /= p.tx;
        const int y0 = (r % p.ty) * TY; r /= p.ty;
        const int z0 = (r % p.tz) * TZ; r /= p.tz;
        const int b = r;
        for (int c = 0; c < p.cin_chunks; ++c, ++phase) {
          if (k3D) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              mbar_wait(&a_empty[j], (fills[j] & 1) ^ 1);
              mbar_expect_tx(&a_full[j], SLOT_TX);
              tma_load_5d(smem + j * SLOT_BYTES, &tmA, &a_full[j], (c & 1) * CT_BLOCK_K, x0 - 1, y0 - 1, z0 - 1 + j,
                          b + p.blkmap[c >> 1] * p.B);
              ++fills[j];
            }
          } else {
            const int j = phase & 1;
            mbar_wait(&a_empty[j], (fills[j] & 1) ^ 1);
            mbar_expect_tx(&a_full[j], SLOT_TX);
            tma_load_5d(smem + j * SLOT_BYTES, &tmA, &a_full[j], (c & 1) * CT_BLOCK_K, x0 - 1, y0 - 1, 0, b + p.blkmap[c >> 1] * p.B);
            ++fills[j];
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================ weight producer ================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
        for (int c = 0; c < p.cin_chunks; ++c)
          for (int t = 0; t < ntaps; ++t, ++it) {
            const uint32_t s = it % C2_BSTAGES, ph = (it / C2_BSTAGES) & 1;
            mbar_wait(&b_empty[s], ph ^ 1);
            mbar_expect_tx(&b_full[s], B_BYTES);
            tma_load_2d(sB + s * CT_B_BYTES, &tmB, &b_full[s], (t * p.cin_chunks + c) * CT_BLOCK_K, 0);
          }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // warp-uniform copy of the TMEM base (redux.sync writes a uniform register): with the plain shared-memory load in the
    // D address the compiler cannot prove uniformity and wraps EVERY tcgen05.mma in an ELECT / R2UR.BROADCAST loop
    const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem_base);
    if (DFL_MMA_ISSUE == 1 || lane == 0) {   // style 1: the whole warp runs the loop, one elected lane issues
      constexpr uint32_t idesc = umma_idesc_bf16(CT_BLOCK_M, kN, 0, 0);
      const uint32_t brick = smem_u32(smem);
      uint32_t fills[NSLOT];
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) fills[s] = 0;
      uint32_t it = 0, tcount = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(&tempty_bar[acc], aph ^ 1);
        tc_fence_after();
        // ncu showed the kernel bound by THIS thread's instruction stream (~90 issue cycles per MMA against the tensor
        // pipe's 64): the D address is built from the uniform TMEM base and the descriptors are advanced by adds
        const uint32_t d_tmem = tmem_u + acc * 256;
        for (int c = 0; c < p.cin_chunks; ++c, ++phase) {
          for (int dz = 0; dz < p.kd; ++dz) {
            // wait for the brick slots this tap group reads
            if (k3D) {
              if (dz == 0) { mbar_wait(&a_full[0], fills[0] & 1); ++fills[0]; mbar_wait(&a_full[1], fills[1] & 1); ++fills[1]; }
              else if (dz == 1) { mbar_wait(&a_full[2], fills[2] & 1); ++fills[2]; }
              else { mbar_wait(&a_full[3], fills[3] & 1); ++fills[3]; }
            } else {
              // slot (phase & 1) is filled every second phase: its fill count so far is phase >> 1.  (A dynamically indexed
              // fills[] array here made the compiler treat every MMA operand as non-uniform: 5 R2UR.BROADCAST per MMA.)
              mbar_wait(&a_full[phase & 1], (phase >> 1) & 1);
            }
            tc_fence_after();
            for (int dy = 0; dy < 3; ++dy)
              for (int dx = 0; dx < 3; ++dx, ++it) {
                const uint32_t s = it % C2_BSTAGES, ph = (it / C2_BSTAGES) & 1;
                mbar_wait(&b_full[s], ph);
                tc_fence_after();
                const uint32_t sb = smem_u32(sB + s * CT_B_BYTES);
                const bool first = (c == 0 && dz == 0 && dy == 0 && dx == 0);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint32_t a0 = k3D ? brick + (h + dz) * SLOT_BYTES + (dy * HX + dx) * 128
                                          : brick + (phase & 1) * SLOT_BYTES + (dy * HX + dx + 8 * h) * 128;
#pragma unroll
                  const uint64_t da0 = umma_desc_sw128(a0, 16, HX * 128), db0 = umma_desc_sw128(sb, 16, 1024);
#pragma unroll
                  for (int k = 0; k < CT_BLOCK_K / 16; ++k) {
                    // K step = +32 B inside the 128-byte swizzle row = +2 in the descriptor's 16-byte address field
                    const uint64_t da = DFL_DESC_INC ? da0 + 2 * k : umma_desc_sw128(a0 + k * 32, 16, HX * 128);
                    const uint64_t db = DFL_DESC_INC ? db0 + 2 * k : umma_desc_sw128(sb + k * 32, 16, 1024);
                    if (DFL_MMA_ISSUE == 0 || elect_one_sync()) umma_bf16(d_tmem + h * CT_BLOCK_N, da, db, idesc, (first && k == 0) ? 0u : 1u);
                  }
                }
                if (DFL_MMA_ISSUE == 0 || elect_one_sync()) umma_commit(&b_empty[s]);
              }
            // release the brick slots no later tap group of this phase reads
            if (k3D) {
              if (DFL_MMA_ISSUE == 0 || elect_one_sync()) umma_commit(&a_empty[dz]);
              if (dz == 2) if (DFL_MMA_ISSUE == 0 || elect_one_sync()) umma_commit(&a_empty[3]);
            } else {
              if (DFL_MMA_ISSUE == 0 || elect_one_sync()) umma_commit(&a_empty[phase & 1]);
            }
          }
        }
        if (DFL_MMA_ISSUE == 0 || elect_one_sync()) umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    // ================================ epilogue (warps 3..6) ================================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int line = row >> 3, xi = row & 7;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tcount) {
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      int r = tile;
      const int x0 = (r % p.tx) * TX; r /= p.tx;
      const int y0 = (r % p.ty) * TY; r /= p.ty;
      const int z0 = (r % p.tz) * TZ; r /= p.tz;
      const int b = r;
      if (kN == CT_BLOCK_N) {
        const EpiWarpT<k3D> T(p, lane, b, k3D ? z0 : 0, y0 + quarter * 4, x0);
        constexpr int NSUB = c2_epi_warps(k3D, kN) / 4;       // warps per TMEM quarter; warp (w - 3) >> 2 takes chunks i % NSUB
        const int esub = (warp - 3) >> 2;
        EpiChunk pre;
        T.prefetch(esub, pre);        // first chunk's mask / residual pieces: in flight while the tile's MMAs finish
        mbar_wait(&tfull_bar[acc], aph);
        tc_fence_after();
        conv_epilogue_warp_t<k3D>(p, T, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 256,
                                  smem_u32(sE) + (warp - 3) * 4096, lane, s_bias, pre, esub, NSUB);
      } else {
        mbar_wait(&tfull_bar[acc], aph);
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int z = k3D ? z0 + h : 0, y = y0 + line, x = k3D ? x0 + xi : x0 + 8 * h + xi;
          const bool valid = (x < p.W) && (y < p.H) && (z < p.D);
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 256 + h * CT_BLOCK_N;
          uint32_t rr[32];
          tmem_ld_32x32(taddr, rr);      // columns >= 16 are never written: ignored
          tmem_ld_wait();
          if (valid) {
            float* o = p.out_f32 + (((static_cast<size_t>(b) * p.D + z) * p.H + y) * p.W + x) * p.cout_small;
            for (int c = 0; c < p.cout_small; ++c) o[c] = __uint_as_float(rr[c]) + s_bias[c];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
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

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static void pick_brick(int D, int H, int W, int& bd, int& bh, int& bw) {
  // 128 voxels; prefer wide-x bricks (contiguous rows), never larger than the (power-of-two-rounded) extent
  auto p2 = [](int v) { int r = 1; while (r < v) r <<= 1; return r; };
  bw = std::min(p2(W), 16);
  bh = std::min(p2(H), 128 / bw);
  bd = 128 / (bw * bh);
  if (D == 1) {           // 2D: spend the remaining factor on x, then y
    while (bd > 1) { if (bw < 128) bw <<= 1; else bh <<= 1; bd >>= 1; }
  } else if (bd > p2(D)) {
    while (bd > p2(D)) { bw <<= 1; bd >>= 1; }
  }
}

// activation tensor map: channels-last blocks [nblk*B, D, H, W, min(cin,128)]
static int make_act_map(CUtensorMap* tm, const void* x, int cin, int nbatch, int D, int H, int W, const uint32_t* box,
                        const uint32_t* estride) {
  const int cb = std::min(cin, 128);
  const uint64_t gd[5] = {static_cast<uint64_t>(cb), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                          static_cast<uint64_t>(D), static_cast<uint64_t>(nbatch)};
  const uint64_t gs[4] = {static_cast<uint64_t>(cb) * 2, static_cast<uint64_t>(cb) * 2 * W,
                          static_cast<uint64_t>(cb) * 2 * W * H, static_cast<uint64_t>(cb) * 2 * W * H * D};
  return encode_tensor_map(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x, gd, gs, box, CU_TENSOR_MAP_SWIZZLE_128B, estride);
}

// Tap-window kernel: 3x3(x3), stride 1, SAME.  cin = 64, 128 or a multiple of 128 (channel-blocked input
// [cin/128 * B, D, H, W, 128]); cout = 128, or 1..16 for the output conv.
int conv_tc_launch(const void* x, const void* w_packed, const float* bias, void* out, void* out2,
                   const void* residual, const void* mask_src, const int64_t* dims /*B,D,H,W*/, int nd, int cin,
                   int cout, int flags, const int32_t* blkmap /*cin/128 entries or null*/, int nphys, cudaStream_t st) {
  const bool small = cout < 128;   // 128 -> 1..3 output conv: w_packed is [16][taps*cin], out is fp32 [.., cout]
  DFL_REQUIRE(cout == 128 || (cout >= 1 && cout <= 16), "conv_tc: Cout must be 128 or <= 16 (got %d)", cout);
  DFL_REQUIRE(cin == 64 || (cin >= 128 && cin % 128 == 0), "conv_tc: Cin must be 64 or a multiple of 128 (got %d)", cin);
  DFL_REQUIRE(nd == 2 || nd == 3, "conv_tc: ndim must be 2 or 3");
  ConvTcParams p{};
  p.B = static_cast<int>(dims[0]);
  p.D = nd == 3 ? static_cast<int>(dims[1]) : 1;
  p.H = static_cast<int>(dims[nd - 1]);
  p.W = static_cast<int>(dims[nd]);
  p.oD = p.D; p.oH = p.H; p.oW = p.W;
  p.kd = nd == 3 ? 3 : 1;
  p.kh = 3;
  p.kw = 3;
  p.bd = nd == 3 ? 2 : 1; p.bh = 16; p.bw = nd == 3 ? 8 : 16;    // tiles: 3D 2x16x8, 2D 1x16x16
  p.tx = (p.W + p.bw - 1) / p.bw;
  p.ty = (p.H + p.bh - 1) / p.bh;
  p.tz = (p.D + p.bd - 1) / p.bd;
  p.ntiles = p.B * p.tz * p.ty * p.tx;
  p.cin_chunks = cin / 64;
  p.flags = flags;
  p.bias = bias;
  p.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.out = small ? nullptr : static_cast<__nv_bfloat16*>(out);
  p.out2 = static_cast<__nv_bfloat16*>(out2);
  p.out_f32 = small ? static_cast<float*>(out) : nullptr;
  p.cout_small = small ? cout : 0;
  DFL_REQUIRE(out || out2, "conv_tc: no output buffer given");
  const int nblk = std::max(1, cin / 128);
  DFL_REQUIRE(nblk <= 8, "conv_tc: at most 8 channel blocks (Cin <= 1024)");
  for (int i = 0; i < 8; ++i) p.blkmap[i] = (blkmap && i < nblk) ? blkmap[i] : i;
  if (nphys <= 0) nphys = nblk;

  CUtensorMap tmA, tmB;
  {
    const uint32_t box[5] = {64, static_cast<uint32_t>(p.bw + 2), static_cast<uint32_t>(p.bh + 2), 1, 1};
    int rc = make_act_map(&tmA, x, cin, nphys * p.B, p.D, p.H, p.W, box, nullptr);   // one halo'd plane per TMA
    if (rc) return rc;
  }
  {
    const int ntaps = p.kd * p.kh * p.kw;
    const uint64_t gd[2] = {static_cast<uint64_t>(ntaps) * cin, static_cast<uint64_t>(small ? 16 : 128)};
    const uint64_t gs[1] = {static_cast<uint64_t>(ntaps) * cin * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(small ? 16 : 128)};
    int rc = encode_tensor_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_packed, gd, gs, box,
                               CU_TENSOR_MAP_SWIZZLE_128B, nullptr);
    if (rc) return rc;
  }
  const int grid = std::min(p.ntiles, num_sms());
  static bool attr_set = false;
  if (!attr_set) {
    DFL_CUDA_OK(cudaFuncSetAttribute(conv_tc2_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES));
    DFL_CUDA_OK(cudaFuncSetAttribute(conv_tc2_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES));
    DFL_CUDA_OK(cudaFuncSetAttribute(conv_tc2_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES));
    DFL_CUDA_OK(cudaFuncSetAttribute(conv_tc2_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES));
    attr_set = true;
  }
  if (nd == 3 && !small)
    conv_tc2_kernel<true, 128><<<grid, c2_threads(true, 128), C2_SMEM_BYTES, st>>>(tmA, tmB, p);
  else if (nd == 3)
    conv_tc2_kernel<true, 16><<<grid, c2_threads(true, 16), C2_SMEM_BYTES, st>>>(tmA, tmB, p);
  else if (!small)
    conv_tc2_kernel<false, 128><<<grid, c2_threads(false, 128), C2_SMEM_BYTES, st>>>(tmA, tmB, p);
  else
    conv_tc2_kernel<false, 16><<<grid, c2_threads(false, 16), C2_SMEM_BYTES, st>>>(tmA, tmB, p);
  DFL_LAUNCH_OK("conv_tc2_kernel");
  return DFL_OK;
}

// Generic per-tap launcher.  in_dims = {nblk*B, D, H, W} of the (channel-blocked) input tensor, tile_dims = {B, D, H, W}
// of the tile domain, out_dims = {D, H, W} of the output tensor; taps = ntap x {dz, dy, dx, kcol}; w_ld = row length of
// the packed weight matrix (elements); the weight pointer addresses the first of the 128 output-channel rows.
int conv_tap_launch(const void* x, const void* w_packed, const float* bias, void* out, void* out2,
                    const void* residual, const void* mask_src, const int64_t* in_dims, const int64_t* tile_dims,
                    const int64_t* out_dims, int nd, int cin, int in_stride, int ntap, const int32_t* taps,
                    int out_stride, const int32_t* out_off, int64_t w_ld, int flags, cudaStream_t st) {
  DFL_REQUIRE(nd == 2 || nd == 3, "conv_tap: ndim must be 2 or 3");
  DFL_REQUIRE(cin == 64 || (cin >= 128 && cin % 128 == 0), "conv_tap: Cin must be 64 or a multiple of 128 (got %d)", cin);
  DFL_REQUIRE(ntap >= 1 && ntap <= CT_MAX_TAPS, "conv_tap: 1..64 taps (got %d)", ntap);
  DFL_REQUIRE(in_stride == 1 || in_stride == 2, "conv_tap: in_stride must be 1 or 2");
  DFL_REQUIRE(out || out2, "conv_tap: no output buffer given");
  ConvTcParams p{};
  p.B = static_cast<int>(tile_dims[0]);
  p.D = nd == 3 ? static_cast<int>(tile_dims[1]) : 1;
  p.H = static_cast<int>(tile_dims[nd - 1]);
  p.W = static_cast<int>(tile_dims[nd]);
  p.oD = nd == 3 ? static_cast<int>(out_dims[0]) : 1;
  p.oH = static_cast<int>(out_dims[nd - 2]);
  p.oW = static_cast<int>(out_dims[nd - 1]);
  p.kd = nd == 3 ? 3 : 1; p.kh = 3; p.kw = 3;
  pick_brick(p.D, p.H, p.W, p.bd, p.bh, p.bw);
  p.tx = (p.W + p.bw - 1) / p.bw;
  p.ty = (p.H + p.bh - 1) / p.bh;
  p.tz = (p.D + p.bd - 1) / p.bd;
  p.ntiles = p.B * p.tz * p.ty * p.tx;
  p.cin_chunks = cin / 64;
  p.flags = flags;
  p.bias = bias;
  p.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out2 = static_cast<__nv_bfloat16*>(out2);
  p.in_stride = in_stride;
  p.out_stride = out_stride;
  p.orz = nd == 3 ? out_off[0] : 0;
  p.ory = out_off[nd - 2];
  p.orx = out_off[nd - 1];
  for (int i = 0; i < 8; ++i) p.blkmap[i] = i;
  p.ntap = ntap;
  for (int t = 0; t < ntap; ++t) {
    p.tap_dz[t] = taps[4 * t]; p.tap_dy[t] = taps[4 * t + 1]; p.tap_dx[t] = taps[4 * t + 2]; p.tap_col[t] = taps[4 * t + 3];
  }
  CUtensorMap tmA, tmB;
  {
    const int Din = nd == 3 ? static_cast<int>(in_dims[1]) : 1;
    const uint32_t s = static_cast<uint32_t>(in_stride);
    // with a traversal stride s the TMA loads ceil(box/s) elements: box = s * voxels wanted
    const uint32_t box[5] = {64, p.bw * s, p.bh * s, (Din == 1 ? 1u : p.bd * s), 1};
    const uint32_t es[5] = {1, s, s, (Din == 1 ? 1u : s), 1};
    int rc = make_act_map(&tmA, x, cin, static_cast<int>(in_dims[0]), Din, static_cast<int>(in_dims[nd - 1]),
                          static_cast<int>(in_dims[nd]), box, es);
    if (rc) return rc;
  }
  {
    const uint64_t gd[2] = {static_cast<uint64_t>(w_ld), 128};
    const uint64_t gs[1] = {static_cast<uint64_t>(w_ld) * 2};
    const uint32_t box[2] = {64, 128};
    int rc = encode_tensor_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_packed, gd, gs, box,
                               CU_TENSOR_MAP_SWIZZLE_128B, nullptr);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    DFL_CUDA_OK(cudaFuncSetAttribute(conv_tap_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_BYTES));
    DFL_CUDA_OK(cudaFuncSetAttribute(conv_tap_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_BYTES_PAIR));
    attr_set = true;
  }
  static const int pair_mode = getenv("DFL_CONV_TAP_PAIR") ? atoi(getenv("DFL_CONV_TAP_PAIR")) : 1;
  if (pair_mode && p.ntiles >= 2 * num_sms()) {        // at least two bricks per SM: pair them (shared weight tiles)
    const int grid = std::min((p.ntiles + 1) / 2, num_sms());
    conv_tap_kernel<true><<<grid, CT_THREADS, CT_SMEM_BYTES_PAIR, st>>>(tmA, tmB, p);
  } else {
    const int grid = std::min(p.ntiles, num_sms());
    conv_tap_kernel<false><<<grid, CT_THREADS, CT_SMEM_BYTES, st>>>(tmA, tmB, p);
  }
  DFL_LAUNCH_OK("conv_tap_kernel");
  return DFL_OK;
}

}  // namespace dfl
