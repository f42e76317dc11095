// deepfluids_b200 -- backward of the 128 -> C (C = 1..3) output convolution (reference model.py:42,84) on tcgen05.
//
// TF autodiff runs Conv*BackpropInput + Conv*BackpropFilter + BiasAddGrad here; both contractions are tiny in one
// GEMM dimension (C*taps <= 81), so instead of a spatial convolution they are expressed over an im2col tile of the
// C-channel output gradient that 128 builder threads assemble directly in shared memory (rows = voxels, columns =
// k = tap*C + co, 128B-swizzled exactly like a TMA image):
//     G[q][k]     = dOut[q - (tap-1)][co]                       (zero outside the domain)
//     ds[q][ci]   = sum_k G[q][k] * W[tap][ci][co]              dgrad:  D1[M=voxel, N=ci]  = G (K-major A) x W'^T
//     dW[k][ci]   = sum_q s[q][ci] * G[q][k]                    wgrad:  D2[M=ci,    N=k]  += s^T (MN-major A) x G (MN-major B)
//     db[co]      = sum_q dOut[q][co]
// The SAME smem image of G is the K-major A operand of the first GEMM and the MN-major B operand of the second.
// s tiles arrive by TMA (two 64-channel boxes); W' (bf16 [128 ci][128 k]) is resident in smem; D1 is double-buffered in
// TMEM and drained by 4 epilogue warps (store ds and ds * lrelu'(y) in bf16); D2 accumulates over the CTA's whole slab
// and is added to the fp32 gradient with atomics at the end.
// Warps: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2..5 = epilogue, 6..9 = im2col builders.
#include <stdlib.h>

#include "dfl_common.cuh"

namespace dfl {

constexpr int LB_THREADS = 320;
constexpr int LB_OP = 32768;                       // one 128 x 128 bf16 operand image (two 16 KB halves)
constexpr int LB_STAGE_F = 3 * 180 * 3;            // floats: halo'd dOut tile, <= 3 planes x (10 x 18) positions x 3 channels
constexpr int LB_TR = 2 * 128 * 128;                // bytes: bf16(v) and bf16(0.2 v) images of one 64-channel half tile
constexpr int LB_SMEM = 5 * LB_OP + LB_TR + 2 * LB_STAGE_F * 4 + 1024 + 1024;   // W' + 2 G + 2 S + transposition images + 2 dOut stages + ctrl + align slack

struct LastBwdParams {
  int B, D, H, W;
  int ty, tx, ntiles;          // tiles of 1 x 8 x 16 voxels
  const float* dout;           // [B,D,H,W,C] fp32
  const float* w;              // [taps][128][C] fp32 (TF layout)
  const __nv_bfloat16* mask_src;   // y of the layer below (lrelu derivative) or nullptr
  __nv_bfloat16* ds;           // [B,D,H,W,128] or nullptr
  __nv_bfloat16* ds_masked;    // [B,D,H,W,128] or nullptr
  float* dw;                   // [taps][128][C] fp32, accumulated
  float* db;                   // [C] fp32, accumulated
  float* det_partial;          // deterministic mode: per-CTA slots of LC_PART_FLOATS floats instead of atomics into dw / db
  // ---- kFuse (2D, C = 1): dOut = dL/dpsi is COMPUTED by the builder warps from the potential and the target (see below)
  const float* pot;            // [B,H,W,1] fp32 stream function psi (the forward kernel's output)
  const float* xt;             // [B,H,W,2] fp32 target velocity
  float* dpot;                 // optional [B,H,W,1]: dL/dpsi for callers that want it
  float* vel;                  // optional [B,H,W,2]: G_ = curl(psi)
  double* partials;            // 2 per CTA
  unsigned int* ticket;
  float* loss3;
  float c1, c2, w1, w2;        // c1 = w1 * grad_scale / N1, c2 = w2 * grad_scale / N2
  double inv_n1, inv_n2;
};

// ---- 2D loss stencil inside the builder warps (kFuse) ------------------------------------------------------------------
// Same formulation and evaluation order as stencil2d_fused_kernel (dfl_stencil.cu): G = curl(psi) (ops.py:264-274), the
// residuals of the four forward differences of G and of the target (ops.py:205-225 on both, trainer.py:145-147,170-172),
// their signs, dL/dG and dL/dpsi = curl^T(dL/dG) with the replicate-last edge folded into the adjoint.  The tile's im2col
// rows need dL/dpsi on the 10 x 18 halo'd tile; that needs psi and x on a 15 x 23 footprint (2 before, 3 after), which the
// 128 builder threads walk in three passes per stage.  The footprint is 3 floats per position against the 256 bytes per
// voxel of the conv's input, so recomputing the halo (2.7x) costs nothing measurable.
constexpr int LS_FY = 15, LS_FX = 23, LS_N = LS_FY * LS_FX;     // 345 footprint positions
__device__ __forceinline__ float ls_sgn(float v) { return (v > 0.f ? 1.f : 0.f) - (v < 0.f ? 1.f : 0.f); }
__device__ __forceinline__ float ls_wgt(int k, int n) { return (k < 0 || k >= n - 1) ? 0.f : (k == n - 2 ? 2.f : 1.f); }
__device__ __forceinline__ float ls_fold(float g0, float g1, int k, int n) {
  return (k < 0 || k >= n - 1) ? 0.f : (k == n - 2 ? g0 + g1 : g0);
}
__device__ __forceinline__ void ls_bar() {
  __syncwarp();
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

__device__ __forceinline__ uint32_t lb_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int C, bool k3D, bool kFuse = false>
__global__ void __launch_bounds__(LB_THREADS, 1)
lastconv_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmS, LastBwdParams p) {
  static_assert(!kFuse || (C == 1 && !k3D), "the fused loss prologue of this kernel is the 2D (C = 1) path");
  constexpr int NT = k3D ? 27 : 9;
  constexpr int KREAL = NT * C;                   // <= 81
  constexpr int KSTEPS1 = (KREAL + 15) / 16;      // K16 steps of the dgrad GEMM
  constexpr int NCHUNK = (KREAL + 7) / 8;         // 16-byte chunks per im2col row that carry data
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                   // [2 halves][128 rows (ci)][128 B]   K-major, k = tap*C+co
  uint8_t* sG = smem + LB_OP;           // 2 buffers
  uint8_t* sS = smem + 3 * LB_OP;       // 2 buffers
  uint8_t* sT = smem + 5 * LB_OP;                              // epilogue transposition images (epilogue warps only)
  float* sD = reinterpret_cast<float*>(sT + LB_TR);            // 2 x LB_STAGE_F: halo'd dOut tiles (builder warps only)
  uint8_t* ctrl = sT + LB_TR + 2 * LB_STAGE_F * 4;
  uint64_t* s_full = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* s_empty = s_full + 2;
  uint64_t* g_full = s_empty + 2;
  uint64_t* g_empty = g_full + 2;
  uint64_t* d1_full = g_empty + 2;
  uint64_t* d1_empty = d1_full + 2;
  uint64_t* d2_full = d1_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d2_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time: zero both G buffers, write W' (bf16, swizzled K-major image) ----
  for (int i = threadIdx.x; i < 3 * LB_OP / 16; i += LB_THREADS) reinterpret_cast<uint4*>(sW)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * KREAL; i += LB_THREADS) {
    const int ci = i / KREAL, k = i % KREAL;
    const int t = k / C, co = k % C;
    const float v = p.w[(static_cast<size_t>(t) * 128 + ci) * C + co];
    const int half = k >> 6, kk = k & 63;
    const int chunk = (kk >> 3) ^ (ci & 7);
    *reinterpret_cast<__nv_bfloat16*>(sW + half * (LB_OP / 2) + ci * 128 + chunk * 16 + (kk & 7) * 2) = __float2bfloat16_rn(v);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1);
      mbar_init(&g_full[s], 128); mbar_init(&g_empty[s], 1);
      mbar_init(&d1_full[s], 1); mbar_init(&d1_empty[s], 128);
    }
    mbar_init(d2_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();          // generic-proxy smem writes (W', zeros) -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int my_tiles = (p.ntiles > static_cast<int>(blockIdx.x))
                           ? (p.ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                           : 0;

  if (warp == 0) {
    // ================================ TMA producer: s tiles ================================
    if (lane == 0) {
      uint32_t i = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++i) {
        int r = tile;
        const int x0 = (r % p.tx) * 16; r /= p.tx;
        const int y0 = (r % p.ty) * 8; r /= p.ty;
        const int z = r % p.D;
        const int b = r / p.D;
        const uint32_t s = i & 1, ph = (i >> 1) & 1;
        mbar_wait(&s_empty[s], ph ^ 1);
        mbar_expect_tx(&s_full[s], LB_OP);
        tma_load_5d(sS + s * LB_OP, &tmS, &s_full[s], 0, x0, y0, z, b);
        tma_load_5d(sS + s * LB_OP + LB_OP / 2, &tmS, &s_full[s], 64, x0, y0, z, b);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem_base);   // uniform-register copy (see conv_tc2_kernel)
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(128, 128, 0, 0);   // G (K-major) x W' (K-major)
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, 128, 1, 1);   // s^T (MN-major) x G (MN-major)
      const uint32_t w0 = smem_u32(sW);
      for (int i = 0; i < my_tiles; ++i) {
        const uint32_t s = i & 1, ph = (i >> 1) & 1;
        mbar_wait(&g_full[s], ph);
        mbar_wait(&s_full[s], ph);
        mbar_wait(&d1_empty[s], ph ^ 1);
        tc_fence_after();
        const uint32_t g0 = smem_u32(sG + s * LB_OP), s0 = smem_u32(sS + s * LB_OP);
#pragma unroll
        const uint64_t dg1 = umma_desc_sw128(g0, 16, 1024), dw1 = umma_desc_sw128(w0, 16, 1024);
#pragma unroll
        for (int k = 0; k < KSTEPS1; ++k) {
          const uint32_t off16 = ((k >> 2) * (LB_OP / 2) + (k & 3) * 32) >> 4;     // descriptor address units (16 B)
          umma_bf16(tmem_u + s * 128, dg1 + off16, dw1 + off16, idesc1, k != 0 ? 1u : 0u);
        }
        umma_commit(&d1_full[s]);
#pragma unroll
        const uint64_t ds2 = umma_desc_sw128(s0, LB_OP / 2, 1024), dg2 = umma_desc_sw128(g0, LB_OP / 2, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_u + 256, ds2 + k * 128, dg2 + k * 128, idesc2, (i != 0 || k != 0) ? 1u : 0u);
        umma_commit(&g_empty[s]);
        umma_commit(&s_empty[s]);
      }
      umma_commit(d2_full);
    }
  } else if (warp < 6) {
    // ================================ epilogue (warps 2..5) ================================
    // Accumulator rows are transposed through shared memory so that the global stores are full 128-byte lines: written
    // straight from the TMEM-row-owning thread, every STG.128 of a warp touched 32 different lines with 16 bytes each
    // (32 half-filled sectors), and that store pattern -- not HBM, not the tensor pipe -- set the kernel's time (a timing
    // run without the stores: 3.4 -> 1.7 ms at 128^3 x 4, profiles/r01_diag_nostore_c4.json).
    //   phase 1 (thread = TMEM row): 64 channels -> bf16(v) and bf16(0.2 v) images [128 rows][128 B], 16-byte chunks
    //            XOR-swizzled by (row & 7);
    //   phase 2 (thread = 16-byte piece of a row; a warp = 4 complete 128-byte row segments): ds = the first image,
    //            ds_masked = per element the first or second image by the sign of the lrelu output (the same values as
    //            rounding v * lrelu'(y) from fp32).
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int te = (warp - 2) * 32 + lane;            // 0..127
    const int px = te >> 3, piece = te & 7;           // phase 2: x column within the tile, 16-byte piece of the 128-byte segment
    const uint32_t sTa = smem_u32(sT), sTb = sTa + 128 * 128;
    const uint32_t w_off = row * 128, sw_w = row & 7;
    int i = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++i) {
      int r = tile;
      const int x = (r % p.tx) * 16 + px; r /= p.tx;
      const int y0 = (r % p.ty) * 8; r /= p.ty;
      const int z = r % p.D;
      const int b = r / p.D;
      const size_t pos0 = ((static_cast<size_t>(b) * p.D + z) * p.H + y0) * p.W + x;    // row (ly = 0, lx = px) of the tile
      const uint32_t s = i & 1, ph = (i >> 1) & 1;
      // L2 prefetch of the NEXT tile's mask lines (256 lines of 128 bytes; the lanes with piece 0 issue them): the mask pieces
      // below are requested only a transposition ahead of their use, which hides an L2 hit but not a DRAM round trip under
      // load (same finding as in the fused 3D kernel, dfl_lastconv_bwd_fused.cu: fb_prefetch_mask)
      if (p.ds_masked && piece == 0 && tile + static_cast<int>(gridDim.x) < p.ntiles) {
        int rn = tile + static_cast<int>(gridDim.x);
        const int xn = (rn % p.tx) * 16 + px; rn /= p.tx;
        const int yn = (rn % p.ty) * 8; rn /= p.ty;
        const int zn = rn % p.D;
        const int bn = rn / p.D;
        if (xn < p.W) {
          const __nv_bfloat16* m0 = p.mask_src + (((static_cast<size_t>(bn) * p.D + zn) * p.H + yn) * p.W + xn) * 128;
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (yn + it < p.H) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(m0 + static_cast<size_t>(it) * p.W * 128));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(m0 + static_cast<size_t>(it) * p.W * 128 + 64));
            }
        }
      }
      mbar_wait(&d1_full[s], ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + s * 128;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        // mask pieces of this half: requested before the transposition so the round trip overlaps it
        uint4 mv[8];
        if (p.ds_masked && x < p.W) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (y0 + it < p.H)
              mv[it] = __ldg(reinterpret_cast<const uint4*>(p.mask_src + (pos0 + static_cast<size_t>(it) * p.W) * 128 + h * 64 + piece * 8));
        }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t rr[32];
          tmem_ld_32x32(taddr + h * 64 + cc * 32, rr);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t wa[4], wb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float v0 = __uint_as_float(rr[q * 8 + 2 * e]), v1 = __uint_as_float(rr[q * 8 + 2 * e + 1]);
              wa[e] = lb_pack(v0, v1);
              wb[e] = lb_pack(v0 * 0.2f, v1 * 0.2f);
            }
            const uint32_t o = w_off + (((cc * 4 + q) ^ sw_w) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sTa + o), "r"(wa[0]), "r"(wa[1]), "r"(wa[2]), "r"(wa[3]) : "memory");
            if (p.ds_masked)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sTb + o), "r"(wb[0]), "r"(wb[1]), "r"(wb[2]), "r"(wb[3]) : "memory");
          }
        }
        if (h == 1) {
          tc_fence_before();
          mbar_arrive(&d1_empty[s]);               // the tensor core may overwrite this accumulator
        }
        __syncwarp();      // warp-aligned barrier after lane-divergent code
        asm volatile("bar.sync 2, 128;" ::: "memory");   // images complete
        if (x < p.W) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {         // row r = it * 16 + px of the tile: (ly = it, lx = px)
            if (y0 + it >= p.H) continue;
            const int rrow = it * 16 + px;
            const uint32_t o = rrow * 128 + ((piece ^ (rrow & 7)) << 4);
            const size_t off = (pos0 + static_cast<size_t>(it) * p.W) * 128 + h * 64 + piece * 8;
            uint4 va;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(va.x), "=r"(va.y), "=r"(va.z), "=r"(va.w) : "r"(sTa + o));
            if (p.ds) *reinterpret_cast<uint4*>(p.ds + off) = va;
            if (p.ds_masked) {
              uint4 vb;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(vb.x), "=r"(vb.y), "=r"(vb.z), "=r"(vb.w) : "r"(sTb + o));
              const uint32_t mw[4] = {mv[it].x, mv[it].y, mv[it].z, mv[it].w};
              const uint32_t aw[4] = {va.x, va.y, va.z, va.w}, bw[4] = {vb.x, vb.y, vb.z, vb.w};
              uint32_t ow[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                // lrelu'(y) = 1 for y >= 0 (incl. -0), else 0.2 (NaN -> 0.2): same rule as lrelu_grad_from_out
                const bool lo1 = __uint_as_float(mw[e] << 16) >= 0.f, hi1 = __uint_as_float(mw[e] & 0xFFFF0000u) >= 0.f;
                ow[e] = ((lo1 ? aw[e] : bw[e]) & 0x0000FFFFu) | ((hi1 ? aw[e] : bw[e]) & 0xFFFF0000u);
              }
              *reinterpret_cast<uint4*>(p.ds_masked + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
          }
        }
        __syncwarp();      // warp-aligned barrier after lane-divergent code
        asm volatile("bar.sync 3, 128;" ::: "memory");   // images may be overwritten
      }
    }
    // ---- D2 -> dW (fp32 atomics), lane = ci ----
    if (my_tiles > 0) {
      mbar_wait(d2_full, 0);
      tc_fence_after();
      const int ci = row;
#pragma unroll 1
      for (int c0 = 0; c0 < ((KREAL + 31) / 32) * 32; c0 += 32) {
        uint32_t rr[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 256 + c0, rr);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int k = c0 + e;
          if (k < KREAL) {
            if (p.det_partial) p.det_partial[static_cast<size_t>(blockIdx.x) * LC_PART_FLOATS + k * 128 + ci] = __uint_as_float(rr[e]);
            else atomicAdd(p.dw + (static_cast<size_t>(k / C) * 128 + ci) * C + (k % C), __uint_as_float(rr[e]));
          }
        }
      }
    }
  } else {
    // ================================ im2col builders (warps 6..9) ================================
    // Each tile's halo'd dOut neighbourhood (NZ planes x 10 x 18 positions x C floats, zero outside the domain) is staged
    // in shared memory once and every im2col row is assembled from it with immediate-offset LDS.  (Gathering the <= 81
    // values of a row straight from global memory, with a bounds test each, made these four warps' instruction stream the
    // limiter of the whole kernel: ncu 769 M warp-instructions per launch at 128^3 x 4, IPC 1.07, DRAM 40 %.)
    constexpr int NZ = k3D ? 3 : 1;
    const int row = (warp - 6) * 32 + lane;
    const int lx = row & 15, ly = row >> 4;
    float bsum[C];
#pragma unroll
    for (int c = 0; c < C; ++c) bsum[c] = 0.f;
    double acc_l1 = 0.0, acc_j = 0.0;       // kFuse: loss terms of the pixels this CTA owns
    int i = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++i) {
      int r = tile;
      const int x0 = (r % p.tx) * 16; r /= p.tx;
      const int y0 = (r % p.ty) * 8; r /= p.ty;
      const int z = r % p.D;
      const int b = r / p.D;
      const uint32_t s = i & 1, ph = (i >> 1) & 1;
      float* st = sD + s * LB_STAGE_F;
      const float* base = p.dout + (static_cast<size_t>(b) * p.D * p.H * p.W) * C;
      if constexpr (kFuse) {
        // stage buffers hold 2 x 180 floats in 2D; the stencil planes live behind them (2 x LB_STAGE_F floats are reserved)
        float* sP = sD + 2 * 180;
        float* sX = sP + LS_N;
        float* sGv = sX + 2 * LS_N;
        float* sDg = sGv + 2 * LS_N;
        static_assert(2 * 180 + 7 * LS_N <= 2 * LB_STAGE_F, "stencil planes do not fit behind the stage buffers");
        const int H = p.H, W = p.W;
        const size_t img = static_cast<size_t>(b) * H * W;
        // S0: psi and x on the footprint (zero outside the domain)
        for (int pos = row; pos < LS_N; pos += 128) {
          const int j = pos / LS_FX, ii = pos - j * LS_FX;
          const int cx = x0 - 3 + ii, cy = y0 - 3 + j;
          const bool in = cx >= 0 && cx < W && cy >= 0 && cy < H;
          const size_t pix = img + static_cast<size_t>(in ? cy : 0) * W + (in ? cx : 0);
          sP[pos] = in ? __ldg(p.pot + pix) : 0.f;
          sX[2 * pos] = in ? __ldg(p.xt + pix * 2) : 0.f;
          sX[2 * pos + 1] = in ? __ldg(p.xt + pix * 2 + 1) : 0.f;
        }
        ls_bar();
        // S1: G = curl(psi): u = d psi / dy, v = psi[x] - psi[x+1]
        for (int pos = row; pos < LS_N; pos += 128) {
          const int j = pos / LS_FX, ii = pos - j * LS_FX;
          const int cx = x0 - 3 + ii, cy = y0 - 3 + j;
          const bool in = cx >= 0 && cx < W && cy >= 0 && cy < H;
          const int im = max(ii - 1, 0), ip = min(ii + 1, LS_FX - 1), jm = max(j - 1, 0), jp = min(j + 1, LS_FY - 1);
          const int ix_lo = (cx <= W - 2) ? ii : im, ix_hi = (cx <= W - 2) ? ip : ii;
          const int jy_lo = (cy <= H - 2) ? j : jm, jy_hi = (cy <= H - 2) ? jp : j;
          float g0 = 0.f, g1 = 0.f;
          if (in) {
            g0 = sP[jy_hi * LS_FX + ii] - sP[jy_lo * LS_FX + ii];
            g1 = sP[j * LS_FX + ix_lo] - sP[j * LS_FX + ix_hi];
            if (p.vel && ii >= 3 && ii < 19 && j >= 3 && j < 11) {        // the tile's own pixels
              float* vd = p.vel + (img + static_cast<size_t>(cy) * W + cx) * 2;
              vd[0] = g0; vd[1] = g1;
            }
          }
          sGv[2 * pos] = g0;
          sGv[2 * pos + 1] = g1;
        }
        ls_bar();
        // S2: dL/dG and the loss terms of the tile's own pixels
        for (int pos = row; pos < LS_N; pos += 128) {
          const int j = pos / LS_FX, ii = pos - j * LS_FX;
          const int cx = x0 - 3 + ii, cy = y0 - 3 + j;
          const bool in = cx >= 0 && cx < W && cy >= 0 && cy < H;
          const bool own = in && ii >= 3 && ii < 19 && j >= 3 && j < 11;
          const int im = max(ii - 1, 0), ip = min(ii + 1, LS_FX - 1), jm = max(j - 1, 0), jp = min(j + 1, LS_FY - 1);
          const float wxm = ls_wgt(cx - 1, W), wx0 = ls_wgt(cx, W), wym = ls_wgt(cy - 1, H), wy0 = ls_wgt(cy, H);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float g = sGv[2 * pos + c], xx = sX[2 * pos + c];
            const float e = g - xx;
            const float dxp = (sGv[2 * (j * LS_FX + ip) + c] - g) - (sX[2 * (j * LS_FX + ip) + c] - xx);
            const float dxm = (g - sGv[2 * (j * LS_FX + im) + c]) - (xx - sX[2 * (j * LS_FX + im) + c]);
            const float dyp = (sGv[2 * (jp * LS_FX + ii) + c] - g) - (sX[2 * (jp * LS_FX + ii) + c] - xx);
            const float dym = (g - sGv[2 * (jm * LS_FX + ii) + c]) - (xx - sX[2 * (jm * LS_FX + ii) + c]);
            const float dg = p.c1 * ls_sgn(e) +
                             p.c2 * ((wxm * ls_sgn(dxm) - wx0 * ls_sgn(dxp)) + (wym * ls_sgn(dym) - wy0 * ls_sgn(dyp)));
            sDg[2 * pos + c] = in ? dg : 0.f;
            if (own) {
              acc_l1 += fabsf(e);
              acc_j += static_cast<double>(wx0 * fabsf(dxp) + wy0 * fabsf(dyp));
            }
          }
        }
        ls_bar();
        // S3: dL/dpsi = D_y^T gU - D_x^T gV on the halo'd tile -> the staged dOut positions (zero outside the domain)
        for (int pos = row; pos < 180; pos += 128) {
          const int yy = pos / 18, xx_ = pos - yy * 18;
          const int j = yy + 2, ii = xx_ + 2;
          const int cx = x0 - 1 + xx_, cy = y0 - 1 + yy;
          const bool in = cx >= 0 && cx < W && cy >= 0 && cy < H;
          float v = 0.f;
          if (in) {
            const int q = j * LS_FX + ii;
            const float dyT_U = ls_fold(sDg[2 * (q - LS_FX)], sDg[2 * q], cy - 1, H) - ls_fold(sDg[2 * q], sDg[2 * (q + LS_FX)], cy, H);
            const float dxT_V = ls_fold(sDg[2 * (q - 1) + 1], sDg[2 * q + 1], cx - 1, W) - ls_fold(sDg[2 * q + 1], sDg[2 * (q + 1) + 1], cx, W);
            v = dyT_U - dxT_V;
            if (p.dpot && xx_ >= 1 && xx_ < 17 && yy >= 1 && yy < 9) p.dpot[img + static_cast<size_t>(cy) * W + cx] = v;
          }
          st[pos] = v;
        }
      } else
      // ---- stage: position pos = (plane, yy, xx) of the halo'd tile <- dOut[z-1+plane (3D), y0-1+yy, x0-1+xx]
      for (int pos = row; pos < NZ * 180; pos += 128) {
        const int pl = pos / 180, rem = pos - pl * 180;
        const int yy = rem / 18, xx = rem - yy * 18;
        const int gz = k3D ? z - 1 + pl : z, gy = y0 - 1 + yy, gx = x0 - 1 + xx;
        const bool inb = gx >= 0 && gx < p.W && gy >= 0 && gy < p.H && gz >= 0 && gz < p.D;
        const float* src = base + ((static_cast<size_t>(inb ? gz : 0) * p.H + (inb ? gy : 0)) * p.W + (inb ? gx : 0)) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) st[pos * C + c] = inb ? __ldg(src + c) : 0.f;
      }
      __syncwarp();      // warp-aligned barrier after lane-divergent code
      asm volatile("bar.sync 1, 128;" ::: "memory");     // stage complete (double-buffered: one barrier per tile suffices)
      mbar_wait(&g_empty[s], ph ^ 1);
      uint8_t* grow = sG + s * LB_OP + row * 128;
      // G[q][k = tap*C+co] = dOut[q - (tap-1)][co]: staged position (2-dz, ly+2-dy, lx+2-dx) relative to the tile origin
      const float* sq = st + (ly * 18 + lx) * C;
#pragma unroll
      for (int j = 0; j < NCHUNK; ++j) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = j * 8 + e;
          v[e] = 0.f;
          if (k < KREAL) {
            const int t = k / C, co = k % C;
            const int dx = t % 3, dy = (t / 3) % 3, dz = k3D ? t / 9 : 0;
            v[e] = sq[(((k3D ? 2 - dz : 0) * 10 + (2 - dy)) * 18 + (2 - dx)) * C + co];
          }
        }
        const int half = j >> 3, jj = j & 7;
        *reinterpret_cast<uint4*>(grow + half * (LB_OP / 2) + ((jj ^ (row & 7)) * 16)) =
            make_uint4(lb_pack(v[0], v[1]), lb_pack(v[2], v[3]), lb_pack(v[4], v[5]), lb_pack(v[6], v[7]));
      }
      // bias gradient: the tile's own voxels = staged centre positions (zero outside the domain)
#pragma unroll
      for (int c = 0; c < C; ++c) bsum[c] += sq[(((k3D ? 1 : 0) * 10 + 1) * 18 + 1) * C + c];
      fence_proxy_async();
      mbar_arrive(&g_full[s]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float t = warp_sum(bsum[c]);
      if (lane == 0 && my_tiles > 0) {
        if (p.det_partial) p.det_partial[static_cast<size_t>(blockIdx.x) * LC_PART_FLOATS + LC_PART_DB + (warp - 6) * 4 + c] = t;
        else atomicAdd(p.db + c, t);
      }
    }
    if constexpr (kFuse) {
      // loss: one fp64 pair per CTA; the last CTA (atomic ticket) adds all pairs in CTA order -> deterministic, no finalize
      // launch (same scheme as lastconv_bwd_fused_kernel)
      double* sred = reinterpret_cast<double*>(sD + 2 * 180);      // the stencil planes are dead
      ls_bar();
      acc_l1 = warp_sum(acc_l1);
      acc_j = warp_sum(acc_j);
      if (lane == 0) { sred[warp - 6] = acc_l1; sred[4 + warp - 6] = acc_j; }
      ls_bar();
      if (row == 0) {
        double a = 0, c = 0;
        for (int q = 0; q < 4; ++q) { a += sred[q]; c += sred[4 + q]; }
        p.partials[2 * blockIdx.x] = a;
        p.partials[2 * blockIdx.x + 1] = c;
        __threadfence();
        const unsigned int done = atomicAdd(p.ticket, 1u);
        if (done == gridDim.x - 1) {
          __threadfence();
          double sa = 0, sc = 0;
          const volatile double* pp = p.partials;
          for (unsigned int q = 0; q < gridDim.x; ++q) { sa += pp[2 * q]; sc += pp[2 * q + 1]; }
          const double l1 = sa * p.inv_n1, jl = sc * p.inv_n2;
          p.loss3[0] = static_cast<float>(p.w1 * l1 + p.w2 * jl);
          p.loss3[1] = static_cast<float>(l1);
          p.loss3[2] = static_cast<float>(jl);
          *p.ticket = 0u;
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

template <int C, bool k3D, bool kFuse = false>
static int lastconv_bwd_launch_t(const CUtensorMap& tmS, const LastBwdParams& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DFL_CUDA_OK(cudaFuncSetAttribute(lastconv_bwd_tc_kernel<C, k3D, kFuse>, cudaFuncAttributeMaxDynamicSharedMemorySize, LB_SMEM));
    attr_set = true;
  }
  const int grid = std::min(p.ntiles, num_sms());
  LastBwdParams q = p;
  size_t wb = 0;
  q.det_partial = deterministic_workspace(&wb);
  DFL_REQUIRE(!q.det_partial || static_cast<size_t>(grid) * LC_PART_FLOATS * sizeof(float) <= wb,
              "lastconv_bwd: deterministic workspace too small");
  lastconv_bwd_tc_kernel<C, k3D, kFuse><<<grid, LB_THREADS, LB_SMEM, st>>>(tmS, q);
  DFL_LAUNCH_OK(kFuse ? "lastconv_bwd_tc_kernel (fused 2D loss)" : "lastconv_bwd_tc_kernel");
  if (q.det_partial) return lastconv_grad_reduce(q.det_partial, grid, q.dw, q.db, (k3D ? 27 : 9) * C, C, st);
  return DFL_OK;
}

// deterministic mode, second pass for both output-conv backward kernels: dw / db += the CTA slots in CTA order
__global__ void __launch_bounds__(128) lastconv_grad_reduce_kernel(const float* __restrict__ partial, int ncta, float* __restrict__ dw,
                                                                   float* __restrict__ db, int kreal, int C) {
  const int k = blockIdx.x, ci = threadIdx.x;
  if (k < kreal) {
    float a = 0.f;
    for (int c = 0; c < ncta; ++c) a += partial[static_cast<size_t>(c) * LC_PART_FLOATS + k * 128 + ci];
    dw[(static_cast<size_t>(k / C) * 128 + ci) * C + (k % C)] += a;
  } else if (ci < C) {
    float a = 0.f;
    for (int c = 0; c < ncta; ++c)
      for (int w = 0; w < 4; ++w) a += partial[static_cast<size_t>(c) * LC_PART_FLOATS + LC_PART_DB + w * 4 + ci];
    db[ci] += a;
  }
}
int lastconv_grad_reduce(const float* partial, int ncta, float* dw, float* db, int kreal, int C, cudaStream_t st) {
  lastconv_grad_reduce_kernel<<<kreal + 1, 128, 0, st>>>(partial, ncta, dw, db, kreal, C);
  DFL_LAUNCH_OK("lastconv_grad_reduce_kernel");
  return DFL_OK;
}

static int lastconv_bwd_map(CUtensorMap* tmS, const void* s, const LastBwdParams& p) {
  const uint64_t gd[5] = {128, static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H), static_cast<uint64_t>(p.D),
                          static_cast<uint64_t>(p.B)};
  const uint64_t gs[4] = {256, 256ull * p.W, 256ull * p.W * p.H, 256ull * p.W * p.H * p.D};
  const uint32_t box[5] = {64, 16, 8, 1, 1};
  return encode_tensor_map(tmS, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, s, gd, gs, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// 2D half of dfl_lastconv_curl_loss_bwd: pot fp32 [B,H,W,1], x fp32 [B,H,W,2], s bf16 [B,H,W,128]; workspace as for 3D
int lastconv_curl_loss_bwd_2d(const void* s, const float* pot, const float* x, const float* w, const void* mask_src, void* ds,
                              void* ds_masked, float* dw, float* db, float* dpot, float* vel, float* loss3, void* workspace,
                              const int64_t* dims, float w1, float w2, float grad_scale, cudaStream_t st) {
  DFL_REQUIRE(s && pot && x && w && dw && db && loss3 && workspace, "lastconv_curl_loss_bwd: null tensor");
  DFL_REQUIRE(!(ds_masked && !mask_src), "lastconv_curl_loss_bwd: ds_masked requested without mask_src");
  LastBwdParams p{};
  p.B = static_cast<int>(dims[0]); p.D = 1; p.H = static_cast<int>(dims[1]); p.W = static_cast<int>(dims[2]);
  DFL_REQUIRE(p.H >= 2 && p.W >= 2, "lastconv_curl_loss_bwd: extents >= 2 required (got %d x %d)", p.H, p.W);
  p.tx = (p.W + 15) / 16;
  p.ty = (p.H + 7) / 8;
  p.ntiles = p.B * p.ty * p.tx;
  p.w = w;
  p.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
  p.ds = static_cast<__nv_bfloat16*>(ds);
  p.ds_masked = static_cast<__nv_bfloat16*>(ds_masked);
  p.dw = dw; p.db = db;
  p.pot = pot; p.xt = x; p.dpot = dpot; p.vel = vel;
  p.partials = static_cast<double*>(workspace);
  p.ticket = reinterpret_cast<unsigned int*>(p.partials + 2 * 160);
  p.loss3 = loss3;
  const double pix = static_cast<double>(p.B) * p.H * p.W, n1 = pix * 2, n2 = pix * 4;
  p.c1 = static_cast<float>(static_cast<double>(w1) * grad_scale / n1);
  p.c2 = static_cast<float>(static_cast<double>(w2) * grad_scale / n2);
  p.w1 = w1; p.w2 = w2;
  p.inv_n1 = 1.0 / n1; p.inv_n2 = 1.0 / n2;
  CUtensorMap tmS;
  int rc = lastconv_bwd_map(&tmS, s, p);
  if (rc) return rc;
  return lastconv_bwd_launch_t<1, false, true>(tmS, p, st);
}

int lastconv_bwd_tc(const void* s, const float* dout, const float* w, const void* mask_src, void* ds, void* ds_masked,
                    float* dw, float* db, const int64_t* dims, int nd, int cout, cudaStream_t st) {
  DFL_REQUIRE(nd == 2 || nd == 3, "lastconv_bwd: ndim must be 2 or 3");
  DFL_REQUIRE(cout >= 1 && cout <= 3, "lastconv_bwd: Cout must be 1..3 (got %d)", cout);
  DFL_REQUIRE(!(ds_masked && !mask_src), "lastconv_bwd: ds_masked requested without mask_src");
  LastBwdParams p{};
  p.B = static_cast<int>(dims[0]);
  p.D = nd == 3 ? static_cast<int>(dims[1]) : 1;
  p.H = static_cast<int>(dims[nd - 1]);
  p.W = static_cast<int>(dims[nd]);
  p.tx = (p.W + 15) / 16;
  p.ty = (p.H + 7) / 8;
  p.ntiles = p.B * p.D * p.ty * p.tx;
  p.dout = dout;
  p.w = w;
  p.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
  p.ds = static_cast<__nv_bfloat16*>(ds);
  p.ds_masked = static_cast<__nv_bfloat16*>(ds_masked);
  p.dw = dw;
  p.db = db;
  CUtensorMap tmS;
  const uint64_t gd[5] = {128, static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H), static_cast<uint64_t>(p.D),
                          static_cast<uint64_t>(p.B)};
  const uint64_t gs[4] = {256, 256ull * p.W, 256ull * p.W * p.H, 256ull * p.W * p.H * p.D};
  const uint32_t box[5] = {64, 16, 8, 1, 1};
  int rc = encode_tensor_map(&tmS, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, s, gd, gs, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (nd == 3) {
    switch (cout) {
      case 1: return lastconv_bwd_launch_t<1, true>(tmS, p, st);
      case 2: return lastconv_bwd_launch_t<2, true>(tmS, p, st);
      default: return lastconv_bwd_launch_t<3, true>(tmS, p, st);
    }
  }
  switch (cout) {
    case 1: return lastconv_bwd_launch_t<1, false>(tmS, p, st);
    case 2: return lastconv_bwd_launch_t<2, false>(tmS, p, st);
    default: return lastconv_bwd_launch_t<3, false>(tmS, p, st);
  }
}

}  // namespace dfl
