// deepfluids_b200 -- forward of the 128 -> C (C = 1..3) output convolution (reference model.py:42,84:
// conv2d/conv3d(x, output_shape[-1], k=last_k, s=1, act=None) through ops.py:12-16) on tcgen05.
//
// The N = 16 variant of the tap-window kernel (dfl_conv_tc.cu) issues one MMA per (tap, K16 step): 432 MMAs per
// 256 voxels, and a tcgen05.mma with M = 128 costs ~64 cycles whatever its N (measured: 4.1 ms per launch at 128^3 x 4
// against an HBM floor of 0.33 ms).  This kernel turns the contraction around, like the fused backward does
// (dfl_lastconv_tc.cu): ONE plain GEMM per input plane, with NO spatial shift on the tensor core,
//     P[u][k] = sum_ci s[u][ci] * W[tap][ci][co]        k = tap*C + co  (<= 81 columns),  u = voxel of the halo'd plane tile
// followed by a shift-sum on the CUDA cores,
//     out[q][co] = bias[co] + sum_tap P[q + tap - 1][tap*C + co]                       (TF SAME: TMA zero fill)
// i.e. 16 MMAs (N = 96) per 128 output voxels instead of 216.  CTAs march along z: every input plane tile (16 x 8 voxels
// + 1 halo = 180 rows) is loaded ONCE and contributes to the three output planes z-1, z, z+1, whose partial sums live in
// registers of the thread that owns the output voxel.
//   smem:  3 stages x [2 x 64-channel halves][184 rows x 128 B] (TMA, 128B swizzle = K-major A operand),
//          W' bf16 [2 halves][NP rows (k)][128 B] (K-major B operand, built once per CTA from the fp32 TF weights),
//          P fp32: a ring of 2 chunks of [32 k][192 rows]  (k-major: both the row-wise writes and the shifted row-wise reads
//          are conflict-free)
//   TMEM:  2 buffers x 2 M-tiles (rows 0..127, 128..255 of the stage) x 128 columns
//   warps: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 4..7 = TMEM -> P (M-tile 0), 8..9 = TMEM -> P (rows 128..179 of
//          M-tile 1), 10..13 = shift-sum + store (one thread per output voxel).
//   The copy warps and the gather warps are a producer / consumer pair over the chunk ring (named barriers FULL[slot]:
//   copy arrives, gather syncs; EMPTY[slot]: gather arrives, copy syncs): the 32-column chunks of plane i+1 are copied while
//   plane i is still being summed.  (Round 1's version ran copy -> barrier -> gather -> barrier on ONE P image with four of the
//   six warps doing both jobs: ~3000 clocks per plane against ~810 clocks of shared-memory instructions and an HBM floor of
//   ~1500.  Measured at 4 x 128^3 inside the step: 0.88 ms -> 0.77 ms with 2 input stages + 3 chunks -> 0.74 ms with 3 input
//   stages + 2 chunks (two stages = 92 KB in flight per SM were TMA-latency-bound).)
#include "dfl_common.cuh"

namespace dfl {

constexpr int LF_THREADS = 448;
constexpr int LF_NST = 3;                  // input stages (two were latency-bound: 92 KB in flight per SM)
constexpr int LF_ROWS = 180;               // 10 x 18 halo'd rows per plane tile
constexpr int LF_HALF = 23552;             // 184 rows x 128 B: one 64-channel half of a stage (1024-aligned)
constexpr int LF_STAGE = 2 * LF_HALF;
constexpr int LF_TX = 2 * LF_ROWS * 128;   // bytes landed per stage fill
constexpr int LF_PR = 192;                 // row pitch of P in floats (multiple of 32: bank = row & 31)
constexpr int LF_EPI = 192;                // copy threads (warps 4..9)
constexpr int LF_GATHER = 128;             // gather threads (warps 10..13)
constexpr int LF_RING = 2;                 // P chunk ring (3 input stages + 2 chunks = 217 KB)
constexpr int LF_CHUNK_F = 32 * LF_PR;     // floats per chunk: 32 k-columns x 192 rows
constexpr int LF_BAR_FULL = 1, LF_BAR_EMPTY = 4;   // named barrier ids 1..3 / 4..6

struct LastFwdParams {
  int B, D, H, W;
  int ty, tx, ncols;           // tile columns of 8 (y) x 16 (x) voxels, marched along z
  const float* w;              // [taps][128][C] fp32 (TF layout)
  const float* bias;           // [C] or nullptr
  float* out;                  // [B,D,H,W,C] fp32
};

template <int C, bool k3D>
struct LFCfg {
  static constexpr int NT = k3D ? 27 : 9;
  static constexpr int KREAL = NT * C;                     // <= 81
  static constexpr int NP = ((KREAL + 15) / 16) * 16;      // MMA N
  static constexpr int W_HALF = NP * 128;                  // bytes per 64-channel half of W'
  static constexpr int NCH = (KREAL + 31) / 32;            // 32-column chunks of P per plane (3D, C = 3: 3)
  static constexpr int SMEM = LF_NST * LF_STAGE + 2 * W_HALF + LF_RING * LF_CHUNK_F * 4 + 1024 /*ctrl*/ + 1024 /*align*/;
};

struct LFSeg { int col, zs, ze; };
// next (tile column, output z-range) segment of this CTA's contiguous share of the (column, plane) list
__device__ __forceinline__ bool lf_next(long long& u, long long u_end, int D, LFSeg& s) {
  if (u >= u_end) return false;
  s.col = static_cast<int>(u / D);
  s.zs = static_cast<int>(u - static_cast<long long>(s.col) * D);
  s.ze = static_cast<int>(min(static_cast<long long>(D), s.zs + (u_end - u)));
  u += s.ze - s.zs;
  return true;
}
// (warp-aligned instructions reached after lane-divergent code: reconverge explicitly, see dfl_lastconv_bwd_fused.cu; every
// bar.sync site below is reached by ONE role only, the other side arrives -- the form compute-sanitizer's synccheck accepts)
__device__ __forceinline__ void lf_bar_sync(int id) {
  __syncwarp();
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(LF_EPI + LF_GATHER) : "memory");
}
__device__ __forceinline__ void lf_bar_arrive(int id) {
  __syncwarp();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(LF_EPI + LF_GATHER) : "memory");
}
// number of input planes this CTA's share touches (the same walk as every role's loop)
__device__ __forceinline__ int lf_count_planes(long long u0, long long u1, int D, bool k3D) {
  int n = 0;
  long long u = u0;
  LFSeg sg;
  while (lf_next(u, u1, D, sg)) {
    const int zlo = k3D ? sg.zs - 1 : sg.zs, zhi = k3D ? sg.ze : sg.ze - 1;
    n += min(zhi, D - 1) - max(zlo, 0) + 1;
  }
  return n;
}

template <int C, bool k3D>
__global__ void __launch_bounds__(LF_THREADS, 1)
lastconv_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmS, LastFwdParams p) {
  using Cfg = LFCfg<C, k3D>;
  constexpr int KREAL = Cfg::KREAL, NP = Cfg::NP, W_HALF = Cfg::W_HALF;
  constexpr int NDZ = k3D ? 3 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem + LF_NST * LF_STAGE;
  constexpr int NCH = Cfg::NCH;
  float* sP = reinterpret_cast<float*>(sW + 2 * W_HALF);
  uint8_t* ctrl = reinterpret_cast<uint8_t*>(sP) + LF_RING * LF_CHUNK_F * 4;
  uint64_t* full = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty = full + LF_NST;
  uint64_t* d_full = empty + LF_NST;
  uint64_t* d_empty = d_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time: W' = bf16 [half][k][64 ci] as a 128B-swizzled K-major image (rows k >= KREAL zero) ----
  for (int i = threadIdx.x; i < 2 * W_HALF / 16; i += LF_THREADS) reinterpret_cast<uint4*>(sW)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * KREAL; i += LF_THREADS) {
    const int tci = i / C, co = i - tci * C;        // i = (t*128 + ci)*C + co: coalesced read of the TF layout
    const int t = tci >> 7, ci = tci & 127;
    const int k = t * C + co;
    const int half = ci >> 6, cc = ci & 63;
    const int chunk = (cc >> 3) ^ (k & 7);
    *reinterpret_cast<__nv_bfloat16*>(sW + half * W_HALF + k * 128 + chunk * 16 + (cc & 7) * 2) = __float2bfloat16_rn(p.w[i]);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS);
    for (int s = 0; s < LF_NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&d_full[b], 1); mbar_init(&d_empty[b], LF_EPI); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();          // generic-proxy smem writes (W') -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // even split of the (tile column, plane) list over the CTAs
  const long long total = static_cast<long long>(p.ncols) * p.D;
  const long long u0 = total * blockIdx.x / gridDim.x, u1 = total * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ================================ TMA producer: halo'd plane tiles of s ================================
    if (lane == 0) {
      uint32_t it = 0;
      long long u = u0;
      LFSeg sg;
      while (lf_next(u, u1, p.D, sg)) {
        int r = sg.col;
        const int x0 = (r % p.tx) * 16; r /= p.tx;
        const int y0 = (r % p.ty) * 8;
        const int b = r / p.ty;
        const int zlo = k3D ? sg.zs - 1 : sg.zs, zhi = k3D ? sg.ze : sg.ze - 1;
        for (int zi = zlo; zi <= zhi; ++zi) {
          if (zi < 0 || zi >= p.D) continue;
          const uint32_t s = it % LF_NST, ph = (it / LF_NST) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], LF_TX);
          tma_load_5d(smem + s * LF_STAGE, &tmS, &full[s], 0, x0 - 1, y0 - 1, zi, b);
          tma_load_5d(smem + s * LF_STAGE + LF_HALF, &tmS, &full[s], 64, x0 - 1, y0 - 1, zi, b);
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer: P = S (K-major A) x W'^T (K-major B) ================================
    const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem_base);   // uniform-register copy (see conv_tc2_kernel)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, NP, 0, 0);
      const uint32_t w0 = smem_u32(sW), a_base = smem_u32(smem);
      uint32_t it = 0;
      long long u = u0;
      LFSeg sg;
      while (lf_next(u, u1, p.D, sg)) {
        const int zlo = k3D ? sg.zs - 1 : sg.zs, zhi = k3D ? sg.ze : sg.ze - 1;
        for (int zi = zlo; zi <= zhi; ++zi) {
          if (zi < 0 || zi >= p.D) continue;
          const uint32_t s = it % LF_NST, ph = (it / LF_NST) & 1, b = it & 1, bph = (it >> 1) & 1;
          mbar_wait(&d_empty[b], bph ^ 1);
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a0 = a_base + s * LF_STAGE;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // M-tile 0 = rows 0..127, M-tile 1 = rows 128..255 of the half (rows >= 180 are never read back)
            const uint64_t da0 = umma_desc_sw128(a0 + h * LF_HALF, 16, 1024);
            const uint64_t da1 = umma_desc_sw128(a0 + h * LF_HALF + 128 * 128, 16, 1024);
            const uint64_t db0 = umma_desc_sw128(w0 + h * W_HALF, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {       // K step = +32 B inside the 128-byte swizzle row = +2 address units
              const uint32_t acc = (h | k) ? 1u : 0u;
              umma_bf16(tmem_u + b * 256, da0 + 2 * k, db0 + 2 * k, idesc, acc);
              umma_bf16(tmem_u + b * 256 + 128, da1 + 2 * k, db0 + 2 * k, idesc, acc);
            }
          }
          umma_commit(&empty[s]);
          umma_commit(&d_full[b]);
          ++it;
        }
      }
    }
  } else if (warp >= 4 && warp < 10) {
    // ================================ copy warps: TMEM -> P chunk ring ================================
    const int mt = (warp - 4) >> 2;                   // M-tile this warp drains: warps 4..7 -> 0, warps 8,9 -> 1
    const int quarter = warp & 3;                     // TMEM lane quarter a warp may access = warp id % 4
    const int prow = mt * 128 + quarter * 32 + lane;  // row of the halo'd plane tile (10 x 18)
    const int nplanes = lf_count_planes(u0, u1, p.D, k3D);
    for (int it = 0; it < nplanes; ++it) {
      const uint32_t bb = it & 1, bph = (it >> 1) & 1;
      mbar_wait(&d_full[bb], bph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + bb * 256 + mt * 128;
#pragma unroll
      for (int q = 0; q < NCH; ++q) {
        const int cc = it * NCH + q, slot = cc % LF_RING;          // chunk counter, ring slot
        uint32_t rr[32];
        tmem_ld_32x32(taddr + q * 32, rr);                        // columns >= NP are never written: ignored
        if (cc >= LF_RING) lf_bar_sync(LF_BAR_EMPTY + slot);      // the gather warps are done with this slot
        tmem_ld_wait();
        if (prow < LF_ROWS) {
          float* dst = sP + slot * LF_CHUNK_F + prow;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (q * 32 + j < KREAL) dst[j * LF_PR] = __uint_as_float(rr[j]);
        }
        lf_bar_arrive(LF_BAR_FULL + slot);
      }
      tc_fence_before();
      mbar_arrive(&d_empty[bb]);                      // the tensor core may overwrite this TMEM buffer
    }
  } else if (warp >= 10) {
    // ================================ gather warps: shift-sum over the chunk ring -> out ================================
    const int ot = (warp - 10) * 32 + lane, lx = ot & 15, ly = ot >> 4;      // output voxel of this thread
    const float* gp = sP + ly * 18 + lx;
    const int total_chunks = lf_count_planes(u0, u1, p.D, k3D) * NCH;
    float bias[C];
#pragma unroll
    for (int c = 0; c < C; ++c) bias[c] = p.bias ? __ldg(p.bias + c) : 0.f;
    int cc = 0;                                       // chunk counter (same sequence as the copy warps')
    long long u = u0;
    LFSeg sg;
    while (lf_next(u, u1, p.D, sg)) {
      int r = sg.col;
      const int x = (r % p.tx) * 16 + lx; r /= p.tx;
      const int y = (r % p.ty) * 8 + ly;
      const int b = r / p.ty;
      const bool valid = (x < p.W) && (y < p.H);
      float* obase = p.out + ((static_cast<size_t>(b) * p.D * p.H + (valid ? y : 0)) * p.W + (valid ? x : 0)) * C;
      const size_t plane = static_cast<size_t>(p.H) * p.W * C;
      float a0[C], a1[C];                             // partial sums of out[zi-1] and out[zi]
#pragma unroll
      for (int c = 0; c < C; ++c) a0[c] = a1[c] = 0.f;
      const int zlo = k3D ? sg.zs - 1 : sg.zs, zhi = k3D ? sg.ze : sg.ze - 1;
      for (int zi = zlo; zi <= zhi; ++zi) {
        float ps[NDZ][C];                             // this plane's contribution per dz
#pragma unroll
        for (int dz = 0; dz < NDZ; ++dz)
#pragma unroll
          for (int c = 0; c < C; ++c) ps[dz][c] = 0.f;
        if (zi >= 0 && zi < p.D) {
#pragma unroll
          for (int q = 0; q < NCH; ++q, ++cc) {
            const int slot = cc % LF_RING;
            lf_bar_sync(LF_BAR_FULL + slot);          // chunk q of this plane has landed
            const float* gq = gp + slot * LF_CHUNK_F;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int k = q * 32 + j;               // k = tap * C + c, tap = (dz*3 + dy)*3 + dx   (compile-time)
              if (k < KREAL) {
                const int tap = k / C, c = k - tap * C;
                const int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;
                ps[dz][c] += gq[j * LF_PR + dy * 18 + dx];
              }
            }
            if (cc + LF_RING < total_chunks) lf_bar_arrive(LF_BAR_EMPTY + slot);   // (nobody waits for the last LF_RING)
          }
        }
        if (k3D) {
          // input plane zi: tap dz = 2 completes out[zi-1], dz = 1 adds to out[zi], dz = 0 opens out[zi+1]
          const int zo = zi - 1;
          if (valid && zo >= sg.zs && zo < sg.ze) {
            float* o = obase + static_cast<size_t>(zo) * plane;
#pragma unroll
            for (int c = 0; c < C; ++c) o[c] = (a0[c] + ps[NDZ - 1][c]) + bias[c];
          }
#pragma unroll
          for (int c = 0; c < C; ++c) { a0[c] = a1[c] + ps[NDZ > 1 ? 1 : 0][c]; a1[c] = ps[0][c]; }
        } else if (valid) {
          float* o = obase + static_cast<size_t>(zi) * plane;
#pragma unroll
          for (int c = 0; c < C; ++c) o[c] = ps[0][c] + bias[c];
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

template <int C, bool k3D>
static int lastconv_fwd_launch_t(const CUtensorMap& tmS, const LastFwdParams& p, cudaStream_t st) {
  using Cfg = LFCfg<C, k3D>;
  static_assert(Cfg::SMEM <= 227 * 1024, "lastconv_fwd: shared-memory plan exceeds 227 KB");
  static bool attr_set = false;
  if (!attr_set) {
    DFL_CUDA_OK(cudaFuncSetAttribute(lastconv_fwd_tc_kernel<C, k3D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const long long total = static_cast<long long>(p.ncols) * p.D;
  const int grid = static_cast<int>(std::min<long long>(total, num_sms()));
  lastconv_fwd_tc_kernel<C, k3D><<<grid, LF_THREADS, Cfg::SMEM, st>>>(tmS, p);
  DFL_LAUNCH_OK("lastconv_fwd_tc_kernel");
  return DFL_OK;
}

// s: bf16 [B,(D,)H,W,128]; w: fp32 TF layout [3,(3,)3,128,cout]; bias: fp32 [cout] or null; out: fp32 [B,(D,)H,W,cout]
int lastconv_fwd_tc(const void* s, const float* w, const float* bias, float* out, const int64_t* dims, int nd, int cout,
                    cudaStream_t st) {
  DFL_REQUIRE(nd == 2 || nd == 3, "lastconv_fwd: ndim must be 2 or 3");
  DFL_REQUIRE(cout >= 1 && cout <= 3, "lastconv_fwd: Cout must be 1..3 (got %d)", cout);
  DFL_REQUIRE(s && w && out, "lastconv_fwd: null tensor");
  LastFwdParams p{};
  p.B = static_cast<int>(dims[0]);
  p.D = nd == 3 ? static_cast<int>(dims[1]) : 1;
  p.H = static_cast<int>(dims[nd - 1]);
  p.W = static_cast<int>(dims[nd]);
  p.tx = (p.W + 15) / 16;
  p.ty = (p.H + 7) / 8;
  p.ncols = p.B * p.ty * p.tx;
  p.w = w;
  p.bias = bias;
  p.out = out;
  CUtensorMap tmS;
  const uint64_t gd[5] = {128, static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H), static_cast<uint64_t>(p.D),
                          static_cast<uint64_t>(p.B)};
  const uint64_t gs[4] = {256, 256ull * p.W, 256ull * p.W * p.H, 256ull * p.W * p.H * p.D};
  const uint32_t box[5] = {64, 18, 10, 1, 1};
  int rc = encode_tensor_map(&tmS, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, s, gd, gs, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (nd == 3) {
    switch (cout) {
      case 1: return lastconv_fwd_launch_t<1, true>(tmS, p, st);
      case 2: return lastconv_fwd_launch_t<2, true>(tmS, p, st);
      default: return lastconv_fwd_launch_t<3, true>(tmS, p, st);
    }
  }
  switch (cout) {
    case 1: return lastconv_fwd_launch_t<1, false>(tmS, p, st);
    case 2: return lastconv_fwd_launch_t<2, false>(tmS, p, st);
    default: return lastconv_fwd_launch_t<3, false>(tmS, p, st);
  }
}

}  // namespace dfl
