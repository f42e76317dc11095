// deepfluids_b200 -- FUSED first backward kernel of the 3D generator step:
//     curl + Jacobian-L1 loss + their adjoints  (the "S" stencil of SURVEY.md 8a)   IN THE PROLOGUE OF
//     the backward of the 128 -> 3 output convolution (dgrad + wgrad + bias-grad on tcgen05, dfl_lastconv_tc.cu).
//
// What the reference runs here (trainer3.py:16-24,49-51 + TF autodiff of model.py:84): ~150-200 slice / sub / concat / abs /
// mean ops that materialise the 9-channel Jacobians twice, then Conv3DBackpropInput / BackpropFilter / BiasAddGrad.  Un-fused
// (round 1) this was four launches between the two output-conv kernels: stencil3d_lean_kernel (reads A, x; writes dL/dA),
// stencil_finalize_kernel, then lastconv_bwd_tc_kernel re-reading dL/dA with its halo.  Here dL/dA never exists in global
// memory: six extra warps of the backward kernel run the stencil's plane pipeline for the CTA's own (tile column, z range)
// and hand every finished dL/dA plane to the im2col builder warps through a 4-deep shared-memory ring.
//
//   warps  0      TMA producer: s tiles (bf16 [128 ch] x 16 x 8 voxels, two 64-channel boxes)
//          1      TMEM alloc + MMA issuer:  ds = G x W'^T (dgrad),  dW += s^T x G (wgrad; G = im2col of dL/dA)
//          2..5   epilogue: TMEM -> ds, ds * lrelu'(y) (bf16, full-line stores through a transposition image); dW at the end
//          6..9   im2col builders: G rows from the three dL/dA planes z-1, z, z+1 in the ring; bias gradient
//          10..15 STENCIL: A[t] -> G = curl A [t-1] -> F = J(G) - J(x), sgn F [t-2] -> dL/dG [t-3] -> dL/dA = curl^T [t-4]
//                 on the tile's 10 x 20 halo'd footprint (+2 +2 stencil halo = 14 x 24 voxels, two x-voxels per thread),
//                 z-marching with the in-plane neighbours in shared memory and the z neighbours in registers -- the same
//                 formulation, evaluation order and boundary rules as stencil3d_lean_kernel (bit-identical G and signs),
//                 single-buffered planes (reads | barrier | writes | barrier among the six warps only).
// CTAs are persistent and MARCH ALONG z over a contiguous share of the (tile column, plane) list, so a column's stencil
// pipeline is warmed up once per segment (4 planes) and every s tile is still loaded exactly once.  The halo'd footprint is
// recomputed per tile column (2.4x the stencil's arithmetic, hidden under the HBM-bound main loop; A and x are 3-channel
// fp32 tensors = 24 of the kernel's ~1050 bytes per voxel, re-read mostly from L2).
// Loss: per-CTA fp64 partial sums of the voxels the CTA OWNS (tile interior, its own z range); the last CTA to finish
// (atomic ticket) adds them in CTA order and writes loss3 -- deterministic, no finalize launch.
#include <stdlib.h>

#include "dfl_common.cuh"

namespace dfl {

// Measured and NOT adopted (4 x 128^3, same box, tools/fused_ab.sh; baseline 2.05-2.13 ms, all BEFORE the stencil warps'
// spill-free loop and the L2 prefetch of the mask lines went in): mask loads with L1::no_allocate
// 2.34 ms; the builder warps' mask pieces staged by cp.async two rounds ahead 2.08 ms; two register sets of mask pieces per
// store thread 2.40-2.49 ms (ptxas tracks every mask LDG of the loop on ONE scoreboard, so waiting for the older set also
// waits for the newer); TMEM base re-read from shared memory instead of its spill reload 2.05 ms; both image reads of a
// store pass before its first store 2.08 ms.  The long-scoreboard samples ncu shows on the store warps' mask loads are slack,
// not the critical path while the stencil warps are.  Adopted: the stencil warps' spill-free state (below), 2.00-2.02 ms, then
// the L2 prefetch of the next tile's mask lines (fb_prefetch_mask), 1.70 ms.

#ifndef FB_MASK_L2PF
#define FB_MASK_L2PF 1         // store warps: prefetch.global.L2 of the NEXT tile's mask lines (0 = A/B build without it)
#endif

constexpr int FB_THREADS = 512;
constexpr int FB_ST_WARP0 = 10;
constexpr int FB_ST_THREADS = 192;                // warps 10..15
constexpr int FB_BUILD = 128;                     // warps 6..9
constexpr int FB_OP = 32768;                      // one 128 x 128 bf16 operand image (two 16 KB halves)
constexpr int FB_TR = 2 * 128 * 64;               // two alternating bf16(v) images of one 32-channel chunk of a tile
constexpr int FB_RING = 4;                        // dL/dA planes in flight between the stencil and the builders
constexpr int FB_PY = 10, FB_PX = 20;             // plane footprint: rows y0-1 .. y0+8, columns x0-2 .. x0+17
constexpr int FB_PITCH = 80;                      // floats per footprint row in the ring: 60 used; 80 = 16 (mod 32) keeps the
                                                  // builders' half-warps (one tile row each, 3-float lane stride) on disjoint banks
constexpr int FB_PLANE_F = FB_PY * FB_PITCH;      // floats per ring plane
constexpr int FB_TPR = FB_PX / 2 + 2;             // stencil thread columns (voxel pairs): footprint + 2 + 2 halo voxels
constexpr int FB_TR_ROWS = FB_PY + 4;             // stencil thread rows
constexpr int FB_NTHR = FB_TPR * FB_TR_ROWS;      // 168 active stencil threads
static_assert(FB_NTHR <= FB_ST_THREADS, "stencil tile does not fit its warps");
constexpr int FB_PS = FB_ST_THREADS + 2;          // float2 slots per stencil smem plane (+1 slot of x+1 overrun, as the lean kernel)
constexpr int FB_FAM = 3 * FB_PS * 2;             // floats per family (three components)
constexpr int FB_SX0 = 5 * FB_FAM;                // float offset of the scalar Sx family
constexpr int FB_ST_F = FB_SX0 + 3 * FB_PS;       // floats of stencil plane storage (single-buffered)
static_assert(FB_ST_F % 2 == 0, "stencil plane storage is zeroed with float2");
// stencil state slots: float2 [szP 3][dgP 3][(wym, c2wym)], each [FB_ST_THREADS] (conflict-free, thread-private)
constexpr int FB_XT_SLOTS = 7, FB_XT_STRIDE = FB_ST_THREADS * 8;
constexpr int FB_XSTATE = FB_XT_SLOTS * FB_XT_STRIDE;
constexpr int FB_SMEM = 5 * FB_OP + FB_TR + FB_RING * FB_PLANE_F * 4 + FB_ST_F * 4 + 1024 + 1024 + FB_XSTATE;
static_assert(FB_SMEM <= 227 * 1024, "fused backward: shared-memory plan exceeds 227 KB");
enum { FB_FA = 0, FB_FG = 1, FB_FX = 2, FB_FS = 3, FB_FD = 4 };
// named barriers: 0 = __syncthreads, 1 = builders, 2/3 = image READY, 4 = stencil warps, 5..8 ring FULL, 9..12 ring EMPTY,
// 13 = epilogue warps, 14/15 = image FREE
constexpr int FB_BAR_ST = 4, FB_BAR_FULL = 5, FB_BAR_EMPTY = 9;

struct FusedBwdParams {
  int B, D, H, W;
  int ty, tx, ncols;            // tile columns of 8 (y) x 16 (x) voxels, marched along z
  const float* pot;             // [B,D,H,W,3] fp32   network output A (vector potential)
  const float* xt;              // [B,D,H,W,3] fp32   target velocity
  const float* w;               // [27][128][3] fp32  (TF layout)
  const __nv_bfloat16* mask_src;
  __nv_bfloat16* ds;
  __nv_bfloat16* ds_masked;
  float* dw;
  float* db;
  float* dpot;                  // optional: dL/dA of the owned voxels (tests / callers that want it)
  float* vel;                   // optional: G = curl(A) of the owned voxels
  double* partials;             // 2 x gridDim.x
  unsigned int* ticket;
  float* det_partial;          // deterministic mode: per-CTA slots (LC_PART_FLOATS) instead of atomics into dw / db
  float* loss3;
  float c1, c2, w1, w2;
  double inv_n1, inv_n2;
};

struct FBSeg { int col, zs, ze; };
__device__ __forceinline__ bool fb_next(long long& u, long long u_end, int D, FBSeg& s) {
  if (u >= u_end) return false;
  s.col = static_cast<int>(u / D);
  s.zs = static_cast<int>(u - static_cast<long long>(s.col) * D);
  s.ze = static_cast<int>(min(static_cast<long long>(D), s.zs + (u_end - u)));
  u += s.ze - s.zs;
  return true;
}

// bar.sync / bar.arrive are warp-ALIGNED instructions: every lane of the warp must execute them together.  Callers reach them
// after lane-divergent code (edge-of-domain predicates in the store loop, in-domain tests of the stencil), and reconvergence
// is otherwise only the compiler's choice (compute-sanitizer synccheck flagged exactly that on overhanging tiles).
__device__ __forceinline__ void fb_bar_sync(int id, int n) {
  __syncwarp();
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void fb_bar_arrive(int id, int n) {
  __syncwarp();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// Epilogue <-> builder hand-over of the transposition images.  A symmetric `bar.sync 2, 256` reached by the two roles from
// two different places of the role-split code is legal PTX, but compute-sanitizer's synccheck models a bar.sync as ONE
// instruction that all participants must reach ("Barrier error: divergent thread(s) in block"; tools/synccheck_probe.cu
// isolates the pattern: kernels B / D flagged, the arrive / sync pair of kernel C accepted), and routing both roles through
// one non-inlined function cost 0.5 ms per launch (ABI register saves in the store loop).  The hand-over is therefore
// written as producer / consumer pairs, every bar.sync site being reached by one role only:
//   READY[img] (ids 2, 3):   epilogue arrives after writing image img, builders sync before reading it;
//   E-wide     (id 13):      the epilogue warps' own "image complete" barrier;
//   FREE[img]  (ids 14, 15): builders arrive after their row passes of image img, the epilogue syncs two rounds later,
//                            before it overwrites that image (the last two rounds arrive nowhere and are skipped).
constexpr int FB_BAR_READY = 2, FB_BAR_EPI = 13, FB_BAR_FREE = 14;

__device__ __forceinline__ uint32_t fb_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------------------------
// stencil plane pipeline (formulation of dfl_stencil3_lean.cu; see there for the derivation of every constant)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fb_sg1(float v) { return (v > 0.f ? 1.f : 0.f) - (v < 0.f ? 1.f : 0.f); }
__device__ __forceinline__ float2 fb_sg2(float2 v) { return make_float2(fb_sg1(v.x), fb_sg1(v.y)); }
__device__ __forceinline__ float2 fsub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 fadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float fabs2(float2 v) { return fabsf(v.x) + fabsf(v.y); }

// interleaved pair of voxels: r[0] = (u0, v0), r[1] = (w0, u1), r[2] = (v1, w1)
struct FBRaw { float2 r[3]; };
__device__ __forceinline__ float2 fbU(const FBRaw& a) { return make_float2(a.r[0].x, a.r[1].y); }
__device__ __forceinline__ float2 fbV(const FBRaw& a) { return make_float2(a.r[0].y, a.r[2].x); }
__device__ __forceinline__ float2 fbW(const FBRaw& a) { return make_float2(a.r[1].x, a.r[2].y); }
__device__ __forceinline__ float2 fbC(const FBRaw& a, int c) { return c == 0 ? fbU(a) : (c == 1 ? fbV(a) : fbW(a)); }

// running addresses of the stencil pipeline: plane t of this thread's voxel pair
struct FBCur { long long eo; };          // element offset; A[t+1], x[t], G[t-1] and dA[t-4] are addressed relative to it

struct FBThread {
  float wx0, wx1, wy, wym, c2wym, ysgn, xs1;
  int yo8, ycase;
  bool inD, outp, own, lastx, region;
  uint32_t xs;         // byte address of this thread's state slot 0 (FB_XT_SLOTS float2 slots, FB_XT_STRIDE apart)
  int zs, ze;          // planes whose dL/dA this segment produces (clipped to the domain)
  int zo_s, zo_e;      // planes this CTA OWNS (loss terms, optional global outputs)
};

__device__ __forceinline__ float2 fb_lds2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float fb_lds1(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void fb_sts2(uint32_t a, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory"); }
__device__ __forceinline__ void fb_sts1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void fb_ld3(const float* q, FBRaw& a) {
  const float2* q2 = reinterpret_cast<const float2*>(q);
  a.r[0] = __ldg(q2); a.r[1] = __ldg(q2 + 1); a.r[2] = __ldg(q2 + 2);
}
__device__ __forceinline__ void fb_st3(float* q, float2 u, float2 v, float2 w) {
  float2* o = reinterpret_cast<float2*>(q);
  o[0] = make_float2(u.x, v.x); o[1] = make_float2(w.x, u.y); o[2] = make_float2(v.y, w.y);
}
constexpr uint32_t fb_fam8(int fam, int c) { return static_cast<uint32_t>((fam * FB_FAM + c * FB_PS * 2) * 4); }

// One z-iteration t of the pipeline for this thread's voxel pair.  sb / ssx: byte addresses of the thread's slot in the
// (single-buffered) float2 planes / the scalar Sx planes.  Register roles as in the lean kernel (ping-pong, swapped by the
// caller): aO = A[t-1] (receives A[t+1]), aN = A[t]; gO = G[t-2], gN <- G[t-1]; xO = x[t-2] (receives x[t]), xN = x[t-1];
// dO = dL/dG[t-4], dN <- dL/dG[t-3].  The carried state that is read once and written once per iteration -- the weighted z
// signs szP (F_z of plane t-3 for the adjoint at t-2) and the partial gradient dgP (completed next iteration by its -1
// neighbours) -- and the (wym, c2 wym) boundary weights live in thread-private shared-memory slots (T.xs): in registers they
// put the stencil warps over the 128-register cap, and the spill reloads (L1 misses beside the streaming traffic) were a third of
// these warps' time in the ncu source page.  Returns dL/dA[t-4] of the pair in (oU, oV, oW) (valid when the caller's uniform test
// says plane t-4 is produced and T.outp).
__device__ __forceinline__ void fb_iter(const int t, const FusedBwdParams& p, const FBThread& T, const uint32_t sb,
                                        const uint32_t ssx, FBRaw& aO, FBRaw& aN, float2 (&gO)[3], float2 (&gN)[3],
                                        FBRaw& xO, FBRaw& xN, float2 (&dO)[3], float2 (&dN)[3], float2 (&ghzP)[2],
                                        float2& dzu, float2& dzv, float& facc_l1,
                                        float& facc_j, FBCur& cur, const int plane3, float2& oU, float2& oV, float2& oW) {
  const float* const pa = p.pot + (cur.eo + plane3);
  const float* const px = p.xt + cur.eo;
  float* const pv = p.vel + (cur.eo - plane3);
  const int D = p.D, zs = T.zs, ze = T.ze;
  constexpr uint32_t tpr8 = FB_TPR * 8;
  bool didG = false, didS = false, didD = false;
  float sxy[3];
  float2 syv[3];
  if (T.inD) {
    // ---- S1: G[q1] = curl(A)[q1], q1 = t-1 (A[q1]'s in-plane neighbours were staged at the end of the last iteration)
    {
      const int q1 = t - 1;
      if (q1 >= 0 && q1 < D && q1 <= ze) {
        if (q1 <= D - 2) { dzu = fsub(fbU(aN), fbU(aO)); dzv = fsub(fbV(aN), fbV(aO)); }   // else: replicate the last z difference
        const float2 y0 = fb_lds2(sb + fb_fam8(FB_FA, 0) + T.yo8), y1 = fb_lds2(sb + fb_fam8(FB_FA, 1) + T.yo8),
                     y2 = fb_lds2(sb + fb_fam8(FB_FA, 2) + T.yo8);
        const float2 yw = make_float2(y1.x, y2.y), yu = make_float2(y0.x, y1.y);
        const float2 aw = fbW(aO), av = fbV(aO), au = fbU(aO);
        const float nw = T.lastx ? aw.x : fb_lds1(sb + fb_fam8(FB_FA, 1) + 8);        // next thread's w0
        const float nv = T.lastx ? av.x : fb_lds1(sb + fb_fam8(FB_FA, 0) + 12);       // next thread's v0
        const float2 dwdx = make_float2(aw.y - aw.x, T.xs1 * (nw - aw.y));
        const float2 dvdx = make_float2(av.y - av.x, T.xs1 * (nv - av.y));
        const float2 dwy = fsub(yw, aw), duy = fsub(yu, au);
        gN[0] = make_float2(fmaf(T.ysgn, dwy.x, -dzv.x), fmaf(T.ysgn, dwy.y, -dzv.y));     // dwdy - dvdz (ops.py:255)
        gN[1] = fsub(dzu, dwdx);                                                         // dudz - dwdx (ops.py:256)
        gN[2] = make_float2(fmaf(-T.ysgn, duy.x, dvdx.x), fmaf(-T.ysgn, duy.y, dvdx.y));   // dvdx - dudy (ops.py:257)
        didG = true;
        if (p.vel && T.own && q1 >= T.zo_s && q1 < T.zo_e) fb_st3(pv, gN[0], gN[1], gN[2]);
      }
    }
    // A[t-1] is dead now: its registers receive A[t+1]
    if (t + 1 >= 0 && t + 1 < D && t + 1 <= ze + 1) fb_ld3(pa, aO);

    // ---- S3: complete dL/dG[q3], q3 = t-3
    {
      const int q3 = t - 3;
      if (q3 >= 0 && q3 < D && q3 >= zs - 1 && q3 <= ze) {
        const float2 wy2 = fb_lds2(T.xs + 6 * FB_XT_STRIDE);        // (wym, c2 * wym)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float sxn = fb_lds1(ssx + c * FB_PS * 4 - 4);
          const float2 syn = fb_lds2(sb + fb_fam8(FB_FS, c) - tpr8);
          const float2 dg = fb_lds2(T.xs + (3 + c) * FB_XT_STRIDE);
          dN[c].x = fmaf(p.c2, fmaf(wy2.x, syn.x, sxn), dg.x);
          dN[c].y = fmaf(wy2.y, syn.y, dg.y);
        }
        didD = true;
      }
    }

    // ---- S2: residuals, signs, loss at q2 = t-2
    {
      const int q2 = t - 2;
      if (q2 >= 0 && q2 < D && q2 >= zs - 2 && q2 <= ze) {
        const float wz = (q2 >= D - 1) ? 0.f : (q2 == D - 2 ? 2.f : 1.f);
        const bool zin = (q2 >= T.zo_s && q2 < T.zo_e);
        FBRaw xy;                                               // x[q2] one row up, and the next thread's voxel 0
        xy.r[0] = fb_lds2(sb + fb_fam8(FB_FX, 0) + tpr8); xy.r[1] = fb_lds2(sb + fb_fam8(FB_FX, 1) + tpr8);
        xy.r[2] = fb_lds2(sb + fb_fam8(FB_FX, 2) + tpr8);
        const float2 xn01 = fb_lds2(sb + fb_fam8(FB_FX, 0) + 8);
        const float xn2 = fb_lds1(sb + fb_fam8(FB_FX, 1) + 8);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float2 Gy = fb_lds2(sb + fb_fam8(FB_FG, c) + tpr8);
          const float Gx = fb_lds1(sb + fb_fam8(FB_FG, c) + 8);
          const float2 Xy = fbC(xy, c);
          const float Xx = c == 0 ? xn01.x : (c == 1 ? xn01.y : xn2);
          const float2 g = gO[c], x = fbC(xO, c);
          const float2 Fx = make_float2((g.y - g.x) - (x.y - x.x), (Gx - g.y) - (Xx - x.y));
          const float2 Fy = fsub(fsub(Gy, g), fsub(Xy, x));
          const float2 Fz = fsub(fsub(gN[c], g), fsub(fbC(xN, c), x));
          const float2 e = fsub(g, x);
          const float2 sx = fb_sg2(Fx), sy = fb_sg2(Fy), sz = fb_sg2(Fz), se = fb_sg2(e);
          const float2 szw = make_float2(wz * sz.x, wz * sz.y);
          const float wsx0 = T.wx0 * sx.x;
          const float2 szp = fb_lds2(T.xs + c * FB_XT_STRIDE);
          const float a0 = fmaf(-T.wy, sy.x, (szp.x - szw.x) - wsx0);
          const float a1 = fmaf(-T.wy, sy.y, fmaf(-T.wx1, sx.y, (szp.y - szw.y) + wsx0));
          fb_sts2(T.xs + (3 + c) * FB_XT_STRIDE, make_float2(fmaf(p.c2, a0, p.c1 * se.x), fmaf(p.c2, a1, p.c1 * se.y)));
          fb_sts2(T.xs + c * FB_XT_STRIDE, szw);
          sxy[c] = sx.y;
          syv[c] = sy;
          if (zin && T.own) {
            facc_l1 += fabs2(e);
            facc_j += fmaf(T.wx0, fabsf(Fx.x), T.wx1 * fabsf(Fx.y)) + fmaf(T.wy, fabs2(Fy), wz * fabs2(Fz));
          }
        }
        didS = true;
      }
    }
    // x[t-2] is dead: its registers receive x[t]
    if (t >= 0 && t < D && t <= ze) fb_ld3(px, xO);

    // ---- S4: dL/dA[r4] = curl^T(dL/dG), r4 = t-4
    {
      const int r4 = t - 4;
      float2 ghz0, ghz1;                        // folded z field gh[r4] of the U, V components (uniform case split)
      if (r4 < 0 || r4 >= D - 1) { ghz0 = ghz1 = make_float2(0.f, 0.f); }
      else if (r4 == D - 2) { ghz0 = fadd(dO[0], dN[0]); ghz1 = fadd(dO[1], dN[1]); }
      else { ghz0 = dO[0]; ghz1 = dO[1]; }
      if (r4 >= zs && r4 < ze && T.outp) {
        const float2 dzT_U = fsub(ghzP[0], ghz0), dzT_V = fsub(ghzP[1], ghz1);
        float2 mW = fb_lds2(sb + fb_fam8(FB_FD, 2) - tpr8), mU = fb_lds2(sb + fb_fam8(FB_FD, 0) - tpr8);
        float2 hW = dO[2], hU = dO[0];
        if (T.ycase == 1) { hW = fadd(hW, fb_lds2(sb + fb_fam8(FB_FD, 2) + tpr8)); hU = fadd(hU, fb_lds2(sb + fb_fam8(FB_FD, 0) + tpr8)); }
        if (T.ycase == 2) { mW = fadd(mW, dO[2]); mU = fadd(mU, dO[0]); hW = make_float2(0.f, 0.f); hU = hW; }
        const float2 dyT_W = fsub(mW, hW), dyT_U = fsub(mU, hU);
        const float lW = fb_lds1(sb + fb_fam8(FB_FD, 2) - 4), lV = fb_lds1(sb + fb_fam8(FB_FD, 1) - 4);
        const float h0W = T.lastx ? dO[2].x + dO[2].y : dO[2].x, h0V = T.lastx ? dO[1].x + dO[1].y : dO[1].x;
        const float h1W = T.lastx ? 0.f : dO[2].y, h1V = T.lastx ? 0.f : dO[1].y;
        const float2 dxT_W = make_float2(lW - h0W, h0W - h1W), dxT_V = make_float2(lV - h0V, h0V - h1V);
        oU = fsub(dzT_V, dyT_W);
        oV = fsub(dxT_W, dzT_U);
        oW = fsub(dyT_U, dxT_V);
      }
      ghzP[0] = ghz0; ghzP[1] = ghz1;
    }
  }
  // ---- every read of the planes written last iteration is done: publish this iteration's planes
  fb_bar_sync(FB_BAR_ST, FB_ST_THREADS);
  if (T.inD) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { fb_sts2(sb + fb_fam8(FB_FA, i), aN.r[i]); fb_sts2(sb + fb_fam8(FB_FX, i), xN.r[i]); }
    if (didG) {
#pragma unroll
      for (int c = 0; c < 3; ++c) fb_sts2(sb + fb_fam8(FB_FG, c), gN[c]);
    }
    if (didD) {
#pragma unroll
      for (int c = 0; c < 3; ++c) fb_sts2(sb + fb_fam8(FB_FD, c), dN[c]);
    }
    if (didS) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { fb_sts1(ssx + c * FB_PS * 4, sxy[c]); fb_sts2(sb + fb_fam8(FB_FS, c), syv[c]); }
    }
  }
  cur.eo += plane3;
  fb_bar_sync(FB_BAR_ST, FB_ST_THREADS);
}

// ------------------------------------------------------------------------------------------------------------------
// Phase 2 of the epilogue, shared by the epilogue warps (row passes 0, 1) and the im2col builder warps (passes 2, 3): one
// pass = 32 tile rows x one 32-channel chunk; 4 lanes move one row's 64 bytes (two complete sectors) from the transposition
// image to ds and ds * lrelu'(y).  The kernel is bound by the instruction latency chain of the warps that run this loop
// (one epilogue warp per scheduler), not by HBM or the tensor pipe, so it is split over eight warps.
struct FBStoreLane { int prow, piece; };
__device__ __forceinline__ void fb_load_mask(const FusedBwdParams& p, const FBStoreLane L, size_t t0, int tx0, int y0, int hh,
                                             int it0, uint4 (&m)[2]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int rr_ = (it0 + q) * 32 + L.prow, ly = rr_ >> 4, lx = rr_ & 15;
    if (p.ds_masked && tx0 + lx < p.W && y0 + ly < p.H)
      m[q] = __ldg(reinterpret_cast<const uint4*>(p.mask_src + (t0 + static_cast<size_t>(ly) * p.W + lx) * 128 + hh * 32 + L.piece * 8));
  }
}
// L2 prefetch of the 128-byte mask line that this lane's pieces of rounds hh and hh + 1 live in, one tile (= one plane of the
// tile column) ahead: the register-landed mask loads of the next tile then hit L2 (~800 clocks) instead of DRAM under load
// (~2 600 clocks, longer than a round, which the one-round-ahead request cannot hide).  1.99-2.00 -> 1.70 ms at 4 x 128^3
// (0.67 -> 0.79 of the measured HBM peak), 1.02 -> 0.89 ms at 16 x 64^3 (profiles/r02b_fused_ab.txt, batch ab4).  Order matters:
// the same prefetch was measured BEFORE the stencil warps' spill reloads were removed and did nothing (2.05 vs 2.05 ms) --
// at that time the stencil warps were the critical path and the store warps' mask wait was slack.
__device__ __forceinline__ void fb_prefetch_mask(const FusedBwdParams& p, const FBStoreLane L, size_t t0, int tx0, int y0, int hh,
                                                 int it0) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int rr_ = (it0 + q) * 32 + L.prow, ly = rr_ >> 4, lx = rr_ & 15;
    if (p.ds_masked && L.piece == 0 && tx0 + lx < p.W && y0 + ly < p.H)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.mask_src + (t0 + static_cast<size_t>(ly) * p.W + lx) * 128 + hh * 32));
  }
}
__device__ __forceinline__ void fb_store_passes(const FusedBwdParams& p, const FBStoreLane L, uint32_t sTa, size_t tile0, int tx0,
                                                int y0, int h, int it0, const uint4 (&mv)[2]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int rr_ = (it0 + q) * 32 + L.prow, ly = rr_ >> 4, lx = rr_ & 15;
    if (tx0 + lx >= p.W || y0 + ly >= p.H) continue;
    const uint32_t o = rr_ * 64 + ((L.piece ^ ((rr_ >> 1) & 3)) << 4);
    const size_t off = (tile0 + static_cast<size_t>(ly) * p.W + lx) * 128 + h * 32 + L.piece * 8;
    uint4 va;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(va.x), "=r"(va.y), "=r"(va.z), "=r"(va.w) : "r"(sTa + o));
    if (p.ds) *reinterpret_cast<uint4*>(p.ds + off) = va;
    if (p.ds_masked) {
      // ds * lrelu'(y): 0.2 * ds for y < 0, evaluated on the bf16 image (a second image holding bf16(0.2 v) rounded from
      // fp32 doubled the shared-memory traffic and the length of this loop for a difference of at most one bf16 ulp in
      // the elements whose two roundings disagree).  lrelu'(y) = 1 for y >= 0 (incl. -0), else 0.2 (NaN -> 0.2): the rule of
      // lrelu_grad_from_out.
      const uint32_t aw[4] = {va.x, va.y, va.z, va.w};
      const uint32_t mw[4] = {mv[q].x, mv[q].y, mv[q].z, mv[q].w};
      uint32_t ow[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a0 = __uint_as_float(aw[e] << 16), a1 = __uint_as_float(aw[e] & 0xFFFF0000u);
        const float k0 = __uint_as_float(mw[e] << 16) >= 0.f ? 1.f : 0.2f, k1 = __uint_as_float(mw[e] & 0xFFFF0000u) >= 0.f ? 1.f : 0.2f;
        ow[e] = fb_pack(a0 * k0, a1 * k1);           // (x * 1.0f is exact: unmasked elements pass through unchanged)
      }
      *reinterpret_cast<uint4*>(p.ds_masked + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FB_THREADS, 1)
lastconv_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ FusedBwdParams p) {
  constexpr int C = 3, NT = 27;
  constexpr int KREAL = NT * C;                   // 81
  constexpr int KSTEPS1 = (KREAL + 15) / 16;      // K16 steps of the dgrad GEMM
  constexpr int NCHUNK = (KREAL + 7) / 8;         // 16-byte chunks per im2col row that carry data
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                   // [2 halves][128 rows (ci)][128 B]   K-major, k = tap*C+co
  uint8_t* sG = smem + FB_OP;           // 2 buffers
  uint8_t* sS = smem + 3 * FB_OP;       // 2 buffers
  uint8_t* sT = smem + 5 * FB_OP;                              // epilogue transposition images
  float* sRing = reinterpret_cast<float*>(sT + FB_TR);         // FB_RING dL/dA planes
  float* sSt = sRing + FB_RING * FB_PLANE_F;                   // stencil planes
  uint8_t* ctrl = reinterpret_cast<uint8_t*>(sSt + FB_ST_F);
  uint8_t* sXt = ctrl + 1024;                                  // stencil warps' thread-private state slots (8-byte aligned)
  uint64_t* s_full = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* s_empty = s_full + 2;
  uint64_t* g_full = s_empty + 2;
  uint64_t* g_empty = g_full + 2;
  uint64_t* d1_full = g_empty + 2;
  uint64_t* d1_empty = d1_full + 2;
  uint64_t* d2_full = d1_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d2_full + 1);
  double* sred = reinterpret_cast<double*>(ctrl + 256);         // 2 x 6 doubles (loss reduction of the stencil warps)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time: zero W' and both G buffers, write W' (bf16, swizzled K-major image), zero the stencil planes ----
  for (int i = threadIdx.x; i < 3 * FB_OP / 16; i += FB_THREADS) reinterpret_cast<uint4*>(sW)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * KREAL; i += FB_THREADS) {
    const int ci = i / KREAL, k = i % KREAL;
    const int t = k / C, co = k % C;
    const float v = p.w[(static_cast<size_t>(t) * 128 + ci) * C + co];
    const int half = k >> 6, kk = k & 63;
    const int chunk = (kk >> 3) ^ (ci & 7);
    *reinterpret_cast<__nv_bfloat16*>(sW + half * (FB_OP / 2) + ci * 128 + chunk * 16 + (kk & 7) * 2) = __float2bfloat16_rn(v);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1);
      mbar_init(&g_full[s], FB_BUILD); mbar_init(&g_empty[s], 1);
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

  // even split of the (tile column, plane) list over the persistent CTAs; every role walks the same segment list
  const long long total = static_cast<long long>(p.ncols) * p.D;
  const long long u0 = total * blockIdx.x / gridDim.x, u1 = total * (blockIdx.x + 1) / gridDim.x;
  const int my_tiles = static_cast<int>(u1 - u0);

  if (warp == 0) {
    // ================================ TMA producer: s tiles ================================
    if (lane == 0) {
      uint32_t i = 0;
      long long u = u0;
      FBSeg sg;
      while (fb_next(u, u1, p.D, sg)) {
        int r = sg.col;
        const int x0 = (r % p.tx) * 16; r /= p.tx;
        const int y0 = (r % p.ty) * 8;
        const int b = r / p.ty;
        for (int z = sg.zs; z < sg.ze; ++z, ++i) {
          const uint32_t s = i & 1, ph = (i >> 1) & 1;
          mbar_wait(&s_empty[s], ph ^ 1);
          mbar_expect_tx(&s_full[s], FB_OP);
          tma_load_5d(sS + s * FB_OP, &tmS, &s_full[s], 0, x0, y0, z, b);
          tma_load_5d(sS + s * FB_OP + FB_OP / 2, &tmS, &s_full[s], 64, x0, y0, z, b);
        }
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
        const uint32_t g0 = smem_u32(sG + s * FB_OP), s0 = smem_u32(sS + s * FB_OP);
        const uint64_t dg1 = umma_desc_sw128(g0, 16, 1024), dw1 = umma_desc_sw128(w0, 16, 1024);
#pragma unroll
        for (int k = 0; k < KSTEPS1; ++k) {
          const uint32_t off16 = ((k >> 2) * (FB_OP / 2) + (k & 3) * 32) >> 4;     // descriptor address units (16 B)
          umma_bf16(tmem_u + s * 128, dg1 + off16, dw1 + off16, idesc1, k != 0 ? 1u : 0u);
        }
        umma_commit(&d1_full[s]);
        const uint64_t ds2 = umma_desc_sw128(s0, FB_OP / 2, 1024), dg2 = umma_desc_sw128(g0, FB_OP / 2, 1024);
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
    // TMEM rows leave through a swizzled shared-memory transposition so that global stores are complete 64-byte row
    // chunks (two full sectors, 4 lanes per row), 32 channels at a time:
    //   phase 1 (thread = TMEM row): 32 channels -> bf16(v) image [128 rows][64 B], 16-byte chunks XOR-swizzled by
    //            (row >> 1) & 3; two images alternate, so one named barrier per round suffices;
    //   phase 2 (thread = 16-byte piece of a row): ds = the image, ds_masked = ds or bf16(0.2 * ds) per element by the sign
    //            of the lrelu output.
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int te = (warp - 2) * 32 + lane;            // 0..127
    const FBStoreLane L{te >> 2, te & 3};             // phase 2: row within a pass of 32 rows, 16-byte piece of its 64 bytes
    const uint32_t sT0 = smem_u32(sT);
    const uint32_t w_off = row * 64, sw_w = (row >> 1) & 3;
    int i = 0;
    long long u = u0;
    FBSeg sg;
    const size_t plane_vox = static_cast<size_t>(p.H) * p.W;
    uint4 mvn[2];                                     // mask pieces of the NEXT 32-channel round (one round in flight)
    while (fb_next(u, u1, p.D, sg)) {
      int r = sg.col;
      const int tx0 = (r % p.tx) * 16; r /= p.tx;
      const int y0 = (r % p.ty) * 8;
      const int b = r / p.ty;
      // lrelu-derivative operand (the layer below's output): requested a whole round ahead of its use, so the DRAM round
      // trip hides behind the previous round's TMEM load, transposition and stores
      fb_load_mask(p, L, ((static_cast<size_t>(b) * p.D + sg.zs) * p.H + y0) * p.W + tx0, tx0, y0, 0, 0, mvn);
      for (int z = sg.zs; z < sg.ze; ++z, ++i) {
        const size_t tile0 = ((static_cast<size_t>(b) * p.D + z) * p.H + y0) * p.W + tx0;    // voxel (ly = 0, lx = 0) of the tile
        const uint32_t s = i & 1, ph = (i >> 1) & 1;
        mbar_wait(&d1_full[s], ph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + s * 128;
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
          uint4 mv[2];
          mv[0] = mvn[0]; mv[1] = mvn[1];
#if FB_MASK_L2PF
          if (!(h & 1) && z + 1 < sg.ze) fb_prefetch_mask(p, L, tile0 + plane_vox, tx0, y0, h, 0);
#endif
          if (h < 3) fb_load_mask(p, L, tile0, tx0, y0, h + 1, 0, mvn);
          else if (z + 1 < sg.ze) fb_load_mask(p, L, tile0 + plane_vox, tx0, y0, 0, 0, mvn);
          const uint32_t sTa = sT0 + (h & 1) * (128 * 64);     // alternating images
          if (i > 0 || h >= 2) fb_bar_sync(FB_BAR_FREE + (h & 1), 128 + FB_BUILD);   // the builders are done with this image
          {
            uint32_t rr[32];
            tmem_ld_32x32(taddr + h * 32, rr);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t wa[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                wa[e] = fb_pack(__uint_as_float(rr[q * 8 + 2 * e]), __uint_as_float(rr[q * 8 + 2 * e + 1]));
              const uint32_t o = w_off + ((q ^ sw_w) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sTa + o), "r"(wa[0]), "r"(wa[1]), "r"(wa[2]), "r"(wa[3]) : "memory");
            }
          }
          if (h == 3) {
            tc_fence_before();
            mbar_arrive(&d1_empty[s]);               // the tensor core may overwrite this accumulator
          }
          // image complete: hand it to the builder warps (row passes 2 and 3 of this round) and to the other epilogue
          // warps (passes 0 and 1)
          fb_bar_arrive(FB_BAR_READY + (h & 1), 128 + FB_BUILD);
          fb_bar_sync(FB_BAR_EPI, 128);
          fb_store_passes(p, L, sTa, tile0, tx0, y0, h, 0, mv);
        }
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
    } else if (p.det_partial) {
      // a CTA without tiles (fewer (column, plane) units than CTAs) still owns a slot of the ordered reduction
      for (int k = 0; k < KREAL; ++k) p.det_partial[static_cast<size_t>(blockIdx.x) * LC_PART_FLOATS + k * 128 + row] = 0.f;
    }
  } else if (warp < FB_ST_WARP0) {
    // ================================ im2col builders (warps 6..9) ================================
    // G[q][k = tap*C+co] = dL/dA[q - (tap-1)][co] from the three ring planes z-1, z, z+1 (zero outside the domain: the
    // stencil warps write zeros there).  Plane i of a segment holds z = zs - 1 + i; slots are claimed in production order.
    // They also run HALF of the epilogue's store loop (row passes 2, 3 of every 32-channel round): the im2col tile of tile
    // i+1 is built first (so the tensor core never waits for it), then the warps join the four rounds of tile i.
    const int row = (warp - 6) * 32 + lane;
    const int lx = row & 15, ly = row >> 4;
    const uint32_t ring0 = smem_u32(sRing);
    const FBStoreLane L{row >> 2, row & 3};
    const uint32_t sT0 = smem_u32(sT);
    const size_t plane_vox = static_cast<size_t>(p.H) * p.W;
    float bsum[C];
#pragma unroll
    for (int c = 0; c < C; ++c) bsum[c] = 0.f;
    // ---- build cursor (one tile ahead of the store cursor)
    int ib = 0, jb = 0;
    uint32_t pbase = 0, waited = 0;           // ring planes produced before the build segment / FULL barriers passed so far
    long long ub = u0;
    FBSeg sb_;
    bool have_b = fb_next(ub, u1, p.D, sb_);
    auto build_next = [&]() {
      if (!have_b) return;
      const int n = sb_.ze - sb_.zs, j = jb;
      while (waited < pbase + j + 3) {
        fb_bar_sync(FB_BAR_FULL + (waited & 3), FB_ST_THREADS + FB_BUILD);
        ++waited;
      }
      const uint32_t s = ib & 1, ph = (ib >> 1) & 1;
      mbar_wait(&g_empty[s], ph ^ 1);
      uint8_t* grow = sG + s * FB_OP + row * 128;
      uint32_t pl[3];                           // shared-space byte addresses (explicit ld.shared: no generic loads)
#pragma unroll
      for (int k = 0; k < 3; ++k) pl[k] = ring0 + (((pbase + j + k) & 3) * FB_PLANE_F + ly * FB_PITCH + lx * C) * 4;
#pragma unroll
      for (int jc = 0; jc < NCHUNK; ++jc) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = jc * 8 + e;
          v[e] = 0.f;
          if (k < KREAL) {
            const int t = k / C, co = k % C;
            const int dx = t % 3, dy = (t / 3) % 3, dz = t / 9;
            // source voxel q - (tap - 1): plane 2 - dz, row ly + 2 - dy, column lx + 3 - dx of the footprint
            v[e] = fb_lds1(pl[2 - dz] + (((2 - dy) * FB_PITCH + (3 - dx) * C + co) * 4));
          }
        }
        const int half = jc >> 3, jj = jc & 7;
        *reinterpret_cast<uint4*>(grow + half * (FB_OP / 2) + ((jj ^ (row & 7)) * 16)) =
            make_uint4(fb_pack(v[0], v[1]), fb_pack(v[2], v[3]), fb_pack(v[4], v[5]), fb_pack(v[6], v[7]));
      }
      // bias gradient: the tile's own voxels = centre positions of plane z (zero outside the domain)
#pragma unroll
      for (int c = 0; c < C; ++c) bsum[c] += fb_lds1(pl[1] + ((1 * FB_PITCH + 2 * C + c) * 4));
      fence_proxy_async();
      mbar_arrive(&g_full[s]);
      fb_bar_arrive(FB_BAR_EMPTY + ((pbase + j) & 3), FB_ST_THREADS + FB_BUILD);     // plane z-1 is dead
      ++ib;
      if (++jb == n) {
        // the segment's last two planes (z = ze-1, ze) were only needed by tiles of this segment
        fb_bar_arrive(FB_BAR_EMPTY + ((pbase + n) & 3), FB_ST_THREADS + FB_BUILD);
        fb_bar_arrive(FB_BAR_EMPTY + ((pbase + n + 1) & 3), FB_ST_THREADS + FB_BUILD);
        pbase += n + 2;
        jb = 0;
        have_b = fb_next(ub, u1, p.D, sb_);
      }
    };
    build_next();                               // tile 0
    // ---- store cursor
    long long u = u0;
    FBSeg sg;
    uint4 mvn[2];
    int is = 0;                                 // tiles stored so far
    while (fb_next(u, u1, p.D, sg)) {
      int r = sg.col;
      const int tx0 = (r % p.tx) * 16; r /= p.tx;
      const int y0 = (r % p.ty) * 8;
      const int b = r / p.ty;
      fb_load_mask(p, L, ((static_cast<size_t>(b) * p.D + sg.zs) * p.H + y0) * p.W + tx0, tx0, y0, 0, 2, mvn);
      for (int z = sg.zs; z < sg.ze; ++z, ++is) {
        const size_t tile0 = ((static_cast<size_t>(b) * p.D + z) * p.H + y0) * p.W + tx0;
        build_next();                           // the NEXT tile's im2col operand, before this tile's store rounds
#pragma unroll 1
        for (int h = 0; h < 4; ++h) {
          uint4 mv[2];
          mv[0] = mvn[0]; mv[1] = mvn[1];
#if FB_MASK_L2PF
          if (!(h & 1) && z + 1 < sg.ze) fb_prefetch_mask(p, L, tile0 + plane_vox, tx0, y0, h, 2);
#endif
          if (h < 3) fb_load_mask(p, L, tile0, tx0, y0, h + 1, 2, mvn);
          else if (z + 1 < sg.ze) fb_load_mask(p, L, tile0 + plane_vox, tx0, y0, 0, 2, mvn);
          fb_bar_sync(FB_BAR_READY + (h & 1), 128 + FB_BUILD);
          fb_store_passes(p, L, sT0 + (h & 1) * (128 * 64), tile0, tx0, y0, h, 2, mv);
          if (is + 1 < my_tiles || h < 2) fb_bar_arrive(FB_BAR_FREE + (h & 1), 128 + FB_BUILD);   // (nobody waits for the last two)
        }
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float t = warp_sum(bsum[c]);
      if (lane == 0 && p.det_partial) p.det_partial[static_cast<size_t>(blockIdx.x) * LC_PART_FLOATS + LC_PART_DB + (warp - 6) * 4 + c] = t;
      else if (lane == 0 && my_tiles > 0) atomicAdd(p.db + c, t);
    }
  } else {
    // ================================ stencil warps (10..15) ================================
    const int tid = threadIdx.x - FB_ST_WARP0 * 32;     // 0..191
    const bool act = tid < FB_NTHR;
    const int r = tid / FB_TPR, k = tid - r * FB_TPR;
    const int D = p.D, H = p.H, W = p.W;
    const int plane3 = H * W * 3;
    const uint32_t st0 = smem_u32(sSt);
    const uint32_t slot = st0 + tid * 8, slot_sx = st0 + FB_SX0 * 4 + tid * 4;
    const uint32_t xs0 = smem_u32(sXt) + tid * 8;
    // this thread's 6 floats of a ring plane (footprint rows 0..9 = thread rows 2..11, pairs 0..9 = thread columns 1..10)
    const bool region = act && r >= 2 && r <= FB_TR_ROWS - 3 && k >= 1 && k <= FB_PX / 2;
    const uint32_t ring_off = smem_u32(sRing) + (region ? ((r - 2) * FB_PITCH + (k - 1) * 6) * 4 : 0);
    float facc_l1 = 0.f, facc_j = 0.f;
    uint32_t pprod = 0;                                  // ring planes produced so far

    auto publish = [&](bool valid, float2 oU, float2 oV, float2 oW) {
      // hand one dL/dA plane to the builders: wait until the slot's previous plane has been consumed, write, signal
      if (pprod >= FB_RING) fb_bar_sync(FB_BAR_EMPTY + (pprod & 3), FB_ST_THREADS + FB_BUILD);
      if (region) {
        const uint32_t o = ring_off + (pprod & 3) * (FB_PLANE_F * 4);
        const float2 z0 = make_float2(0.f, 0.f);
        fb_sts2(o, valid ? make_float2(oU.x, oV.x) : z0);
        fb_sts2(o + 8, valid ? make_float2(oW.x, oU.y) : z0);
        fb_sts2(o + 16, valid ? make_float2(oV.y, oW.y) : z0);
      }
      fb_bar_arrive(FB_BAR_FULL + (pprod & 3), FB_ST_THREADS + FB_BUILD);
      ++pprod;
    };

    long long u = u0;
    FBSeg sg;
    while (fb_next(u, u1, p.D, sg)) {
      int rc = sg.col;
      const int x0 = (rc % p.tx) * 16; rc /= p.tx;
      const int y0 = (rc % p.ty) * 8;
      const int b = rc / p.ty;
      FBThread T;
      T.zo_s = sg.zs; T.zo_e = sg.ze;
      T.zs = max(sg.zs - 1, 0); T.ze = min(sg.ze + 1, D);
      const int cy = y0 - 3 + r, cx0 = x0 - 4 + 2 * k;
      T.inD = act && cy >= 0 && cy < H && cx0 >= 0 && cx0 < W;
#ifdef FB_DIAG_NO_STENCIL      // timing diagnostic only (results are garbage): the stencil warps publish zero planes
      T.inD = false;
#endif
      const bool top = (cy == H - 1);
      T.region = region;
      T.outp = T.inD && region;
      T.own = T.inD && r >= 3 && r <= FB_TR_ROWS - 4 && k >= 2 && k <= FB_PX / 2 - 1;      // the 8 x 16 tile interior
      T.lastx = (cx0 == W - 2);
      T.wx0 = T.lastx ? 2.f : 1.f;
      T.wx1 = T.lastx ? 0.f : 1.f;
      T.wy = (cy >= H - 1) ? 0.f : (cy == H - 2 ? 2.f : 1.f);
      T.wym = top ? 2.f : 1.f;
      T.c2wym = p.c2 * T.wym;
      T.xs = xs0;
      fb_sts2(xs0 + 6 * FB_XT_STRIDE, make_float2(T.wym, T.c2wym));
#pragma unroll
      for (int c = 0; c < 6; ++c) fb_sts2(xs0 + c * FB_XT_STRIDE, make_float2(0.f, 0.f));      // szP = dgP = 0
      T.ysgn = top ? -1.f : 1.f;
      T.xs1 = T.lastx ? -1.f : 1.f;
      T.yo8 = top ? -FB_TPR * 8 : FB_TPR * 8;
      T.ycase = (cy == H - 2) ? 1 : (top ? 2 : 0);
      const size_t base = (static_cast<size_t>(b) * D * H + (T.inD ? cy : 0)) * W * 3 + (T.inD ? cx0 : 0) * 3;

      // ---- reset: zero every stencil plane (out-of-domain threads never write; a previous segment must not leak)
      fb_bar_sync(FB_BAR_ST, FB_ST_THREADS);
      for (int q = tid; q < FB_ST_F / 2; q += FB_ST_THREADS) reinterpret_cast<float2*>(sSt)[q] = make_float2(0.f, 0.f);
      fb_bar_sync(FB_BAR_ST, FB_ST_THREADS);

      FBRaw a0, a1, xr0, xr1;
      float2 g0[3], g1[3], d0[3], d1[3], ghzP[2];
      const float2 z2 = make_float2(0.f, 0.f);
      float2 dzu = z2, dzv = z2;
#pragma unroll
      for (int c = 0; c < 3; ++c) a0.r[c] = a1.r[c] = xr0.r[c] = xr1.r[c] = g0[c] = g1[c] = d0[c] = d1[c] = z2;
      ghzP[0] = ghzP[1] = z2;

      // a zero plane stands in for z = -1 (conv padding)
      if (sg.zs == 0) publish(false, z2, z2, z2);

      // iterations t = ts .. te in pairs (ts even-aligned downwards so that the register ping-pong has a fixed phase)
      const int ts = (T.zs - 2) & ~1, te = T.ze + 3;
      if (T.inD) {
        if (ts >= 0 && ts < D) fb_ld3(p.pot + base + static_cast<size_t>(ts) * plane3, a1);
        if (ts - 1 >= 0 && ts - 1 < D) fb_ld3(p.xt + base + static_cast<size_t>(ts - 1) * plane3, xr1);
      }
      // the A plane of iteration ts must be in shared memory before iteration ts + 1 reads its neighbours: the iteration
      // publishes aN itself (end-of-iteration stores), so nothing to do here.
      const long long off = static_cast<long long>(base) + static_cast<long long>(ts) * plane3;
      FBCur cur{off};
      for (int t = ts; t <= te; t += 2) {
        float2 oU = z2, oV = z2, oW = z2;
        fb_iter(t, p, T, slot, slot_sx, a0, a1, g0, g1, xr0, xr1, d0, d1, ghzP, dzu, dzv, facc_l1, facc_j, cur,
                plane3, oU, oV, oW);
        {
          const int r4 = t - 4;
          if (r4 >= T.zs && r4 < T.ze) {                      // uniform
            publish(T.outp, oU, oV, oW);
            if (p.dpot && T.own && r4 >= T.zo_s && r4 < T.zo_e) fb_st3(p.dpot + (cur.eo - 5LL * plane3), oU, oV, oW);   // eo: plane t+1
          }
        }
        oU = oV = oW = z2;
        fb_iter(t + 1, p, T, slot, slot_sx, a1, a0, g1, g0, xr1, xr0, d1, d0, ghzP, dzu, dzv, facc_l1, facc_j, cur,
                plane3, oU, oV, oW);
        {
          const int r4 = t - 3;
          if (r4 >= T.zs && r4 < T.ze) {
            publish(T.outp, oU, oV, oW);
            if (p.dpot && T.own && r4 >= T.zo_s && r4 < T.zo_e) fb_st3(p.dpot + (cur.eo - 5LL * plane3), oU, oV, oW);
          }
        }
      }
      // a zero plane stands in for z = D
      if (sg.ze == D) publish(false, z2, z2, z2);
    }

    // ---- loss: reduce over the stencil warps, one fp64 pair per CTA; the last CTA adds all pairs in CTA order
    const double acc_l1 = warp_sum(static_cast<double>(facc_l1));
    const double acc_j = warp_sum(static_cast<double>(facc_j));
    if (lane == 0) { sred[warp - FB_ST_WARP0] = acc_l1; sred[6 + warp - FB_ST_WARP0] = acc_j; }
    fb_bar_sync(FB_BAR_ST, FB_ST_THREADS);
    if (tid == 0) {
      double a = 0, c = 0;
      for (int q = 0; q < 6; ++q) { a += sred[q]; c += sred[6 + q]; }
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
        *p.ticket = 0u;                                    // ready for the next launch (stream order)
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

size_t lastconv_curl_loss_bwd_workspace_bytes() { return (2 * 160 + 2) * sizeof(double); }

// pot, x: fp32 [B,D,H,W,3]; s: bf16 [B,D,H,W,128] (the output conv's input); see include/deepfluids_b200.h
int lastconv_curl_loss_bwd(const void* s, const float* pot, const float* x, const float* w, const void* mask_src, void* ds,
                           void* ds_masked, float* dw, float* db, float* dpot, float* vel, float* loss3, void* workspace,
                           const int64_t* dims, float w1, float w2, float grad_scale, cudaStream_t st) {
  DFL_REQUIRE(s && pot && x && w && dw && db && loss3 && workspace, "lastconv_curl_loss_bwd: null tensor");
  DFL_REQUIRE(!(ds_masked && !mask_src), "lastconv_curl_loss_bwd: ds_masked requested without mask_src");
  FusedBwdParams p{};
  p.B = static_cast<int>(dims[0]); p.D = static_cast<int>(dims[1]); p.H = static_cast<int>(dims[2]); p.W = static_cast<int>(dims[3]);
  DFL_REQUIRE(p.D >= 2 && p.H >= 2 && p.W >= 2 && (p.W & 1) == 0, "lastconv_curl_loss_bwd: extents >= 2 and even W required "
              "(got %d x %d x %d); use dfl_stencil_loss_fwdbwd + dfl_lastconv_bwd otherwise", p.D, p.H, p.W);
  DFL_REQUIRE(static_cast<long long>(p.H) * p.W * 3 < (1LL << 31), "lastconv_curl_loss_bwd: plane too large");
  p.tx = (p.W + 15) / 16;
  p.ty = (p.H + 7) / 8;
  p.ncols = p.B * p.ty * p.tx;
  p.pot = pot; p.xt = x; p.w = w;
  p.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
  p.ds = static_cast<__nv_bfloat16*>(ds);
  p.ds_masked = static_cast<__nv_bfloat16*>(ds_masked);
  p.dw = dw; p.db = db; p.dpot = dpot; p.vel = vel;
  p.partials = static_cast<double*>(workspace);
  p.ticket = reinterpret_cast<unsigned int*>(p.partials + 2 * 160);
  p.loss3 = loss3;
  const double vox = static_cast<double>(p.B) * p.D * p.H * p.W;
  const double n1 = vox * 3, n2 = vox * 9;
  p.c1 = static_cast<float>(static_cast<double>(w1) * grad_scale / n1);
  p.c2 = static_cast<float>(static_cast<double>(w2) * grad_scale / n2);
  p.w1 = w1; p.w2 = w2;
  p.inv_n1 = 1.0 / n1; p.inv_n2 = 1.0 / n2;
  CUtensorMap tmS;
  const uint64_t gd[5] = {128, static_cast<uint64_t>(p.W), static_cast<uint64_t>(p.H), static_cast<uint64_t>(p.D),
                          static_cast<uint64_t>(p.B)};
  const uint64_t gs[4] = {256, 256ull * p.W, 256ull * p.W * p.H, 256ull * p.W * p.H * p.D};
  const uint32_t box[5] = {64, 16, 8, 1, 1};
  int rc = encode_tensor_map(&tmS, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, s, gd, gs, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    DFL_CUDA_OK(cudaFuncSetAttribute(lastconv_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
    attr_set = true;
  }
  const long long total = static_cast<long long>(p.ncols) * p.D;
  const int grid = static_cast<int>(std::min<long long>(total, std::min(num_sms(), 160)));
  size_t wb = 0;
  p.det_partial = deterministic_workspace(&wb);
  DFL_REQUIRE(!p.det_partial || static_cast<size_t>(grid) * LC_PART_FLOATS * sizeof(float) <= wb,
              "lastconv_curl_loss_bwd: deterministic workspace too small");
  lastconv_bwd_fused_kernel<<<grid, FB_THREADS, FB_SMEM, st>>>(tmS, p);
  DFL_LAUNCH_OK("lastconv_bwd_fused_kernel");
  if (p.det_partial) return lastconv_grad_reduce(p.det_partial, grid, p.dw, p.db, 81, 3, st);
  return DFL_OK;
}

}  // namespace dfl
