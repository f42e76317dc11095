/* deepfluids_b200 -- C ABI of the B200 (sm_100a) hot-path library  (libdeepfluids_b200.so)
 *
 * The reference (byungsook/deep-fluids) has no plugin/FFI layer: its hot path is Python graph code whose
 * arithmetic runs inside TensorFlow 1.15.  This header is the boundary a maintainer would bind instead of
 * those TF ops (see INTEGRATION.md for the ctypes stub); each entry point names the reference interface it
 * replaces.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller (PyTorch allocates);
 *   - channels-last tensors: NHWC [B,H,W,C] / NDHWC [B,D,H,W,C]  (reference ops.py:228 "bzyxd");
 *   - `dims` = {B,H,W} (ndim = 2) or {B,D,H,W} (ndim = 3);  `stream` is a cudaStream_t passed as void*;
 *   - dtype codes: DFL_F32 = 0, DFL_BF16 = 1;
 *   - asynchronous on `stream`, never allocates, never synchronises;
 *   - returns 0 on success, a negative DFL_ERR_* code otherwise; message via dfl_last_error() (thread-local).
 * There is NO CPU fallback: without a CUDA device / sm_100a the calls fail with DFL_ERR_CUDA.
 */
#ifndef DEEPFLUIDS_B200_H_
#define DEEPFLUIDS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFL_F32 0
#define DFL_BF16 1

#define DFL_OK 0
#define DFL_ERR_ARG (-1)
#define DFL_ERR_CUDA (-2)
#define DFL_ERR_UNSUPPORTED (-3)
#define DFL_ERR_INIT (-4)

/* conv epilogue flags */
#define DFL_CONV_LRELU 1          /* y = max(v, 0.2 v)                      (ops.py:9-10)            */
#define DFL_CONV_OUT2_UPSAMPLE 2  /* out2 is written nearest-x2 upsampled   (ops.py:75-91)          */
#define DFL_CONV_MASK_AFTER_RESIDUAL 4 /* out2 = (v + residual) * lrelu'(mask_src) instead of v*lrelu' + residual */
#define DFL_CONV_SPLIT_IO 8       /* fp32-grade mode: bf16 tensors are (hi, lo) pairs one channel block apart      */

/* ---- lifecycle ------------------------------------------------------------------------------------- */
int dfl_version(void);
const char* dfl_last_error(void);
/* bind the calling thread to CUDA device `device`, resolve the TMA driver entry point, query the SM count.
 * Replaces config.py:76-77 (CUDA_VISIBLE_DEVICES selection) for the library. */
int dfl_init(int device);

/* ---- finite-difference stencils (reference ops.py:205-290) ------------------------------------------- */
/* ops.curl (ops.py:264-274; 2D, reads channel 0 of a `pot_channels`-channel tensor) and the 3D curl =
 * second return of ops.jacobian3 (ops.py:255-260, trainer3.py:18).  vel: 2 (2D) / 3 (3D) channels. */
int dfl_curl_fwd(const void* pot, void* vel, const int64_t* dims, int ndim, int pot_channels, int dtype, void* stream);
/* ops.jacobian (ops.py:205-225): jac 4 ch, vort 1 ch;  ops.jacobian3 (ops.py:227-262): jac 9 ch, curl 3 ch.
 * `jac` or `vort_or_curl` may be NULL. */
int dfl_jacobian_fwd(const void* vel, void* jac, void* vort_or_curl, const int64_t* dims, int ndim, int dtype,
                     void* stream);
/* ops.divergence / ops.divergence3 (ops.py:276-290): out [B,(D-1,)H-1,W-1,1]. */
int dfl_divergence(const void* vel, void* div, const int64_t* dims, int ndim, int dtype, void* stream);

/* Adjoints of the three stencils above (fp32): what TF autodiff derives from the slice / sub / concat graphs of
 * ops.py:205-274 when curl / jacobian / jacobian3 sit inside a differentiated graph -- e.g. arch=dg, whose discriminator
 * sees concat(G_, vorticity(G_)) (trainer.py:149-156, trainer3.py:27-34).  The train step of arch=de/ae does not use them
 * (dfl_stencil_loss_fwdbwd emits dL/dpot directly).
 *   dfl_curl_bwd:     dpot [..,dpot_channels] = curl^T(dvel)   (2D: channel 0, the others 0; 3D: dpot_channels = 3)
 *   dfl_jacobian_bwd: dvel [..,ndim] = J^T(djac) + aux^T(daux); djac [..,ndim^2] or NULL, daux [..,1|3] or NULL */
int dfl_curl_bwd(const float* dvel, float* dpot, const int64_t* dims, int ndim, int dpot_channels, void* stream);
int dfl_jacobian_bwd(const float* djac, const float* daux, float* dvel, const int64_t* dims, int ndim, void* stream);
/* LSGAN terms of arch=dg (trainer.py:174-176, trainer3.py:53-55): loss[0] = mean((d - target)^2) over n values;
 * dd (may be NULL) = scale * d loss / d d.  Deterministic single-block reduction. */
int dfl_mse_loss(const float* d, float target, float* loss, float* dd, size_t n, float scale, void* stream);

/* Un-fused L1 term for use_curl=False (trainer.py:141-144: the generator emits the velocity itself, so there is no curl to
 * fuse with): loss[0] = mean|a - b| over n values; dd (may be NULL) = scale * sgn(a - b) / n, added to dd when `accumulate`.
 * The Jacobian term is the same call on dfl_jacobian_fwd outputs, its gradient goes back through dfl_jacobian_bwd.
 * workspace: dfl_l1_loss_workspace_bytes() bytes, zeroed once by the caller; deterministic (ordered fp64 partials). */
size_t dfl_l1_loss_workspace_bytes(void);
int dfl_l1_loss(const float* a, const float* b, float* loss, float* dd, size_t n, float scale, int accumulate,
                void* workspace, void* stream);

/* Fused loss + gradient (SURVEY.md 8a "S").  Replaces, in one pass: curl (trainer.py:140 / trainer3.py:18),
 * jacobian of prediction and target (trainer.py:32,146 / trainer3.py:24), both L1 means (trainer.py:170-172 /
 * trainer3.py:49-51) and the TF autodiff of all of it w.r.t. the network output.
 *   pot   [..,pot_channels] network output (stream function, channel 0 / 3-ch vector potential)
 *   x     [..,2|3]          target velocity
 *   dpot  [..,1|3]          OUT  grad_scale * dL/dpot                    (dtype = dtype_pot)
 *   vel   [..,2|3] or NULL  OUT  G_ = curl(pot)                          (dtype = dtype_pot)
 *   loss3 float[3]          OUT  {w1*l1 + w2*jl1, l1, jl1}  (means; deterministic reduction)
 *   workspace               dfl_stencil_loss_workspace_bytes(dims, ndim) bytes */
size_t dfl_stencil_loss_workspace_bytes(const int64_t* dims, int ndim);
int dfl_stencil_loss_fwdbwd(const void* pot, const void* x, void* dpot, void* vel, float* loss3, void* workspace,
                            const int64_t* dims, int ndim, int pot_channels, float w1, float w2, float grad_scale,
                            int dtype_pot, int dtype_x, void* stream);

/* same; the 2D gradient is written with `dpot_channels` channels (channel 0 = d/d psi, the others 0) so it can feed a
 * network whose output has more than one channel (2D AE with use_curl: trainer.py:359-361 curls channel 0 only). */
int dfl_stencil_loss_fwdbwd_ex(const void* pot, const void* x, void* dpot, void* vel, float* loss3, void* workspace,
                               const int64_t* dims, int ndim, int pot_channels, int dpot_channels, float w1, float w2,
                               float grad_scale, int dtype_pot, int dtype_x, void* stream);

/* ---- fully connected (slim.fully_connected, activation None: ops.py:23-24, model.py:19,61) ------------ */
/* out[b,n] = sum_k z[b,k] W[k,n] + bias[n];  z,W,bias fp32 (W in TF [in,out] layout), K <= 16, B <= 64. */
int dfl_fc_fwd(const float* z, const float* W, const float* bias, void* out, int B, int K, int N, int out_dtype,
               void* stream);
/* dW[k,n] = sum_b z[b,k] dout[b,n];  db[n] = sum_b dout[b,n]   (overwritten) */
int dfl_fc_bwd(const float* z, const void* dout, float* dW, float* db, int B, int K, int N, int dout_dtype,
               void* stream);

/* General fully connected layer (slim.fully_connected at the ops level, ops.py:23-24; the MLP of arch=nn, model.py:218-224):
 * C[M,N] = op(A)[M,K] op(B)[K,N] (+ bias[N]) (+ C if accumulate), fp32.  A row-major [M,K] ([K,M] if transA), B row-major
 * [K,N] ([N,K] if transB).  forward: A = x, B = W;  dx = dy W^T: transB;  dW = x^T dy: transA;  db = dfl_colsum_f32(dy). */
int dfl_gemm_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int transA, int transB,
                 int accumulate, void* stream);
int dfl_colsum_f32(const float* x, float* out, int M, int N, void* stream);

/* ---- 3x3 / 3x3x3 convolution, stride 1, SAME (slim.conv2d / slim.conv3d: ops.py:12-16) ----------------- */
/* fp32 TF-layout weights [taps][Cin][Cout] (HWIO / DHWIO) -> bf16 GEMM operands for forward and dgrad. */
int dfl_pack_conv_weights(const float* w, void* w_fwd, void* w_dgrad, int taps, int cin, int cout, void* stream);
/* The same for n_layers layers of identical shape in one launch: ptr_table = DEVICE array of 3 * n_layers 64-bit addresses,
 * [0,n) the fp32 weights, [n,2n) the forward operands, [2n,3n) the dgrad operands (re-pack after every optimizer step). */
int dfl_pack_conv_weights_multi(const void* ptr_table, int n_layers, int taps, int cin, int cout, void* stream);
/* Implicit-GEMM conv on tcgen05 tensor cores, bf16 in / fp32 accumulate, Cin % 64 == 0, Cout == 128.
 *   v    = conv(x, w_packed) + bias;   if (flags & DFL_CONV_LRELU) v = lrelu(v);
 *   if (mask_src) v *= lrelu'(mask_src)                       (dgrad: derivative of the layer below)
 *   if (out)  out  = v
 *   if (out2) out2 = v + residual (residual may be NULL), nearest-x2 upsampled if DFL_CONV_OUT2_UPSAMPLE
 * Forward use: model.py:26-36 / :68-79 (conv + lrelu [+ residual add + upscale]).  Backward use (dgrad): pass
 * w_dgrad from dfl_pack_conv_weights, bias = NULL.
 * Small-Cout variant (the 128 -> 1..3 output conv, model.py:42,84): cout in [1,16], w_packed = bf16 [16][taps*Cin]
 * (rows >= cout zero), `out` = fp32 [..,cout] = conv + bias; no activation / out2 / residual / mask. */
int dfl_conv3x3_fwd(const void* x, const void* w_packed, const float* bias, void* out, void* out2,
                    const void* residual, const void* mask_src, const int64_t* dims, int ndim, int cin, int cout,
                    int flags, void* stream);
/* dW[tap][ci][co] += sum_p x[p+tap-1][ci] * dpre[p][co]   (fp32, TF layout, accumulated: zero it first);
 * if db != NULL also db[co] += sum_p dpre[p][co] (BiasAddGrad), computed as one more GEMM in a free TMEM slot. */
int dfl_conv3x3_wgrad(const void* x, const void* dpre, float* dw, float* db, const int64_t* dims, int ndim, int cin,
                      int cout, void* stream);
/* db[c] += sum_p dpre[p][c]   (dpre bf16 [npos][128]) */
int dfl_bias_grad(const void* dpre, float* db, size_t npos, void* stream);

/* Forward of the 128 -> cout (1..3) output conv (model.py:42,84: conv2d/conv3d(x, output_shape[-1], k=3, s=1, act=None)):
 * out = conv(s, w) + bias, fp32 [..,cout].  s = bf16 [..,128]; w = the fp32 TF-layout variable [3,(3,)3,128,cout] (rounded
 * to bf16 on chip); bias may be NULL.  One plain GEMM per input plane (P[voxel][tap*cout+co], N <= 96) + a shift-sum over
 * the taps, marching along z: 13x fewer tensor-core instructions than the cout <= 16 variant of dfl_conv3x3_fwd. */
int dfl_lastconv_fwd(const void* s, const float* w, const float* bias, float* out, const int64_t* dims, int ndim, int cout,
                     void* stream);

/* Fused tensor-core backward of the output conv: ds = conv^T(dout, w) (bf16, may be NULL), ds_masked = ds *
 * lrelu'(mask_src) (bf16, may be NULL), dw += s^T (x) dout, db += sum dout (fp32, accumulated).  s = the conv's input
 * (bf16 [..,128]).  Replaces Conv*BackpropInput + Conv*BackpropFilter + BiasAddGrad of model.py:42,84. */
int dfl_lastconv_bwd(const void* s, const float* dout, const float* w, const void* mask_src, void* ds, void* ds_masked,
                     float* dw, float* db, const int64_t* dims, int ndim, int cout, void* stream);

/* FUSED last-generator-epilogue / first-backward-prologue pair of the train step (north_star; SURVEY.md 8a "S" + a3-last); 3D
 * described first, 2D (ndim = 2) at the end.
 * Replaces trainer3.py:16-24,49-51 (curl, jacobian3 of prediction and target, both L1 means) + TF autodiff of all of it +
 * Conv3DBackpropInput / Conv3DBackpropFilter / BiasAddGrad of model.py:84 -- with NOTHING launched in between:
 *   dfl_lastconv_curl_loss_fwd: the output conv; its epilogue adds the bias and stores the fp32 potential A = G_s
 *       (same kernel as dfl_lastconv_fwd -- the curl / loss half needs a +-2 voxel halo of A that a forward tile does not
 *       own, so it lives in the backward kernel's prologue, where only the 3-channel A and x are re-read);
 *   dfl_lastconv_curl_loss_bwd: ONE kernel.  Six stencil warps compute G_ = curl(A), the residuals J(G_) - J(x), their signs,
 *       dL/dG_ and dL/dA = curl^T(dL/dG_) plane by plane (z-march, in-plane neighbours in shared memory) and hand each dL/dA
 *       plane through a shared-memory ring to the im2col builders of the tensor-core backward (ds = conv^T(dL/dA, w), ds_masked
 *       = ds * lrelu'(mask_src), dw += s^T (x) dL/dA, db += sum dL/dA).  dL/dA is never written to global memory unless the
 *       caller passes `dpot` (and `vel` for G_); loss3 = {w1*l1 + w2*jl1, l1, jl1} is complete when the kernel ends
 *       (ordered fp64 partials, last CTA adds them: deterministic).
 *   pot, x: fp32 [B,D,H,W,3];  s: bf16 [B,D,H,W,128];  W even.  workspace: dfl_lastconv_curl_loss_workspace_bytes() bytes,
 *   ZEROED ONCE by the caller before the first launch (it holds the CTA ticket, which resets itself).
 *   ndim = 2 (trainer.py:140-147,170-172 + model.py:42): pot = the stream function psi fp32 [B,H,W,1], x fp32 [B,H,W,2], s bf16
 *   [B,H,W,128], w [3,3,128,1].  The four im2col-builder warps of the output conv's backward compute G_ = curl(psi), the
 *   Jacobian residuals, the loss terms and dL/dpsi on each 8 x 16 tile's 15 x 23 footprint (three barrier-separated passes in
 *   shared memory) instead of loading a dL/dpsi tile: same outputs, same workspace, same deterministic loss reduction. */
size_t dfl_lastconv_curl_loss_workspace_bytes(void);
int dfl_lastconv_curl_loss_fwd(const void* s, const float* w, const float* bias, float* pot, const int64_t* dims, int ndim,
                               int cout, void* stream);
int dfl_lastconv_curl_loss_bwd(const void* s, const float* pot, const float* x, const float* w, const void* mask_src,
                               void* ds, void* ds_masked, float* dw, float* db, float* dpot, float* vel, float* loss3,
                               void* workspace, const int64_t* dims, int ndim, float w1, float w2, float grad_scale,
                               void* stream);

/* adjoint of nearest-x2 upsampling fused with the lrelu derivative (model.py:35-36 / :77-78 backward):
 *   ds = sum of the 2x2(x2) children of g;  dmasked = ds * lrelu'(mask_src).  cdims = COARSE dims. */
int dfl_pool_mask(const void* g, const void* mask_src, void* ds, void* dmasked, const int64_t* cdims, int ndim,
                  void* stream);
/* Operands of the phase-decomposed upsample-conv: conv3(upsample_x2(s)) at the fine voxel 2p + r equals, per phase r (one
 * bit per axis), a convolution with 2 taps per axis on the COARSE tensor whose weights are sums of the layer's taps
 * (r = 0: {w0 | w1 + w2} on offsets {-1, 0}; r = 1: {w0 + w1 | w2} on {0, +1}): 8/27 of the dense FLOPs in 3D, 4/9 in 2D.
 *   w fp32 TF layout [3^nd][cin][cout] -> w_fwd bf16 [P][cout][T*cin], w_dgrad bf16 [cin][P*T*cout], P = T = 2^ndim;
 * run through dfl_conv_taps: forward = P launches (in_stride 2 on the up-sampled tensor = its coarse source, out_stride 2,
 * out_off = phase), data gradient = ONE launch with all P*T taps on the coarse grid. */
int dfl_pack_phase_weights(const float* w, void* w_fwd, void* w_dgrad, int ndim, int cin, int cout, void* stream);
/* Weight gradient of that layer on 8/27 (3D) of the dense FLOPs: dfl_phase_wgrad runs the tensor-core weight-gradient kernel
 * as a 4^ndim-tap stride-2 correlation between dy on the FINE grid ([B,(2D,)2H,2W,128] bf16) and the layer's COARSE input s
 * ([B,(D,)H,W,128] bf16) into t_scratch (fp32 [4^ndim][128][128], zero it first); dfl_phase_wgrad_fold adds the folded result
 * to the TF-layout gradient dw [3^ndim][cin][cout].  The bias gradient is dfl_bias_grad(dy). */
int dfl_phase_wgrad(const void* dy_fine, const void* s_coarse, float* t_scratch, const int64_t* fine_dims,
                    const int64_t* coarse_dims, int ndim, void* stream);
int dfl_phase_wgrad_fold(const float* t_scratch, float* dw, int ndim, int cin, int cout, void* stream);
/* Deterministic weight / bias gradients (new; TF's own GPU kernels for Conv*BackpropFilter are not bitwise reproducible
 * either).  By default the split-K partial sums of dfl_conv3x3_wgrad / dfl_conv_wgrad_ex / dfl_phase_wgrad / dfl_bias_grad
 * meet in dw / db through fp32 red.global.add / atomicAdd, whose order varies from run to run.  After
 * dfl_set_deterministic(workspace, bytes) with a caller-owned device buffer of dfl_deterministic_workspace_bytes() bytes,
 * every CTA stores its partial accumulators to its own slot of the workspace and a second, stream-ordered launch adds them
 * to dw / db in slab order: two runs on the same inputs then give bit-identical gradients.  The launches that use the
 * workspace must not overlap each other (one stream).  dfl_set_deterministic(NULL, 0) restores the atomic reduction.
 * The 128 -> 1..3 output conv's dW / db (dfl_lastconv_bwd, dfl_lastconv_curl_loss_bwd) follow the same switch (per-CTA
 * slots, added in CTA order), and so do dfl_enc_fc_fwd (per-block partial sums, added in block order by a second launch) and
 * dfl_gemm_f32 (no split-K in this mode): whole generator and auto-encoder train steps are reproducible bit for bit. */
size_t dfl_deterministic_workspace_bytes(void);
int dfl_set_deterministic(void* workspace, size_t bytes);
/* coarse[b,(z,)y,x,:] = fine[b,(2z,)2y,2x,:] (bf16, 128 channels): the coarse source s of an up-sampled tensor upscale(s) */
int dfl_gather_stride2(const void* fine, void* coarse, const int64_t* cdims, int ndim, void* stream);
/* same with a coarse-grid addend: ds = sum of the children of g + addend.  Used by the phase-decomposed upsample-conv
 * (model.py:76-79 followed by :67-69 = eight 2x2x2 convolutions on the coarse tensor): the first conv's data gradient lands
 * on the coarse grid directly (`addend`), only the residual branch's gradient `g` still has to be pooled. */
int dfl_pool_mask_add(const void* g, const void* addend, const void* mask_src, void* ds, void* dmasked, const int64_t* cdims,
                      int ndim, void* stream);

/* ---- encoder / auto-encoder extensions (EncoderBE/EncoderBE3, AE/AE3: model.py:118-216; trainer.py:357-396) ------
 * Activations wider than 128 channels are stored as channel blocks [nblk*B,(D,)H,W,128] (block-major), so the concat
 * of model.py:144,180 is free: producers write straight into their block.  dfl_conv3x3_fwd accepts such inputs
 * (cin a multiple of 128). */
/* forward operand with the input-channel count padded to cin_ld (first conv: 2/3 channels zero-padded to 128) */
int dfl_pack_conv_weights_ex(const float* w, void* w_fwd, void* w_dgrad, int taps, int cin, int cout, int cin_ld,
                             void* stream);
/* Generic per-tap tensor-core convolution (one TMA box per tap): stride-2 convolutions (model.py:140,176; TMA element
 * stride 2, TF SAME padding via the tap offsets), their data gradient (one launch per output parity class: tap subset +
 * out_stride/out_off place the results on the fine grid) or any explicit tap list.  in_dims = {nblk*B,(D,)H,W} of the
 * input, tile_dims = {B,(D,)H,W} of the tile domain, out_dims = {(D,)H,W} of the output tensor, taps = ntap x {dz,dy,dx,
 * kcol}, w_ld = row length of the packed weight matrix whose 128 output rows start at w_packed.  Epilogue as
 * dfl_conv3x3_fwd. */
int dfl_conv_taps(const void* x, const void* w_packed, const float* bias, void* out, void* out2, const void* residual,
                  const void* mask_src, const int64_t* in_dims, const int64_t* tile_dims, const int64_t* out_dims, int ndim,
                  int cin, int in_stride, int ntap, const int32_t* taps, int out_stride, const int32_t* out_off,
                  int64_t w_ld, int flags, void* stream);
/* fp32-grade weight gradient on split operands: x2, dpre2 are (hi, lo) bf16 pairs [2B,(D,)H,W,128] (see
 * dfl_split_f32); dw += x_hi^T dP_hi + x_lo^T dP_hi + x_hi^T dP_lo per tap (fp32 accumulate), db += sum(dP_hi + dP_lo).
 * One launch: the three operand combinations are accumulated in the same TMEM tile before the reduction into dw. */
int dfl_conv3x3_wgrad_split(const void* x2, const void* dpre2, float* dw, float* db, const int64_t* dims, int ndim,
                            void* stream);
/* weight gradient of one (128-channel input block, 128-channel output block) pair; in_stride 2 samples x at
 * 2p + tap - pad (stride-2 conv); element (tap, ci, co) is added at dw[tap*dw_tap_stride + ci*dw_row_stride + co]. */
int dfl_conv_wgrad_ex(const void* x, const void* dpre, float* dw, float* db, const int64_t* x_dims, const int64_t* dims,
                      int ndim, int in_stride, int pad, int dw_tap_stride, int dw_row_stride, void* stream);
/* fp32 [n][cin] -> bf16 [n][128] with zero padding (encoder input) */
int dfl_pad_cast(const float* in, void* out, size_t n, int cin, void* stream);
/* ops.upscale / ops.upscale3 (ops.py:66-91) as standalone layers: out[b,(2z+c,)2y+a,2x+e,:] = in[b,(z,)y,x,:] and the adjoint
 * out[b,(z,)y,x,:] = sum of the 2^ndim children of g.  cdims = COARSE {B,(D,)H,W}; any channel count; DFL_F32 or DFL_BF16.
 * (Inside the fused engines the up-sampling is the conv epilogue's replicated store and the adjoint is dfl_pool_mask.) */
int dfl_upscale2(const void* in, void* out, const int64_t* cdims, int ndim, int channels, int dtype, void* stream);
int dfl_pool2(const void* g, void* out, const int64_t* cdims, int ndim, int channels, int dtype, void* stream);
/* out = (a + b) * lrelu'(y);  b, y may be NULL;  bf16, n % 8 == 0 */
int dfl_add_mask(const void* a, const void* b, const void* y, void* out, size_t n, void* stream);
/* encoder FC (model.py:149,185) on the channel-blocked concat tensor [nblk][B][V][128]: z = flat W + bias (Z<=16,B<=8)*/
int dfl_enc_fc_fwd(const void* flat, const float* W, const float* bias, float* z, int B, int V, int nblk, int Z,
                   void* stream);
int dfl_enc_fc_bwd(const void* flat, const float* W, const float* dz, float* dW, float* db, void* dflat, int B, int V,
                   int nblk, int Z, void* stream);
/* decoder FC input gradient: dz[b][k] (+)= sum_n dout[b][n] W[k][n] */
int dfl_fc_dz(const void* dout, const float* W, float* dz, int B, int K, int N, int dout_dtype, int accumulate,
              void* stream);
/* AE parameter loss (trainer.py:385-387): loss_p = mean((y - z[:, Z-P:])^2); dz = scale * d loss_p / dz */
int dfl_ae_loss_p(const float* z, const float* y, float* dz, float* loss_p, int B, int Z, int P, float scale,
                  void* stream);

/* use_sparse (model.py:196,210; trainer.py:389-394): z = sigmoid(z_lin);  backward: dz += w5 * d/dz sum_j KL(Bernoulli(rho)
 * || Bernoulli(mean_b z[:, j])) over the first Z-P dims, then dz_lin = dz * z (1 - z); loss_kl written. */
int dfl_ae_sigmoid(const float* z_lin, float* z, int n, void* stream);
int dfl_ae_sparse_bwd(const float* z, float* dz, float* dz_lin, float* loss_kl, int B, int Z, int P, float rho, float w5,
                      void* stream);

/* ---- fp32-grade mode "bf16x3" (BASELINE config 2: the reference's fp32 arithmetic on bf16 tensor cores) -------------
 * A logical fp32 tensor is a (hi, lo) pair of bf16 tensors (hi = bf16(v), lo = bf16(v - hi)) stored as two consecutive
 * channel blocks [2*B,(D,)H,W,128].  conv(x, w) ~ x_hi*w_hi + x_lo*w_hi + x_hi*w_lo is one dfl_conv3x3_fwd_ex launch with
 * cin = 384 (virtual blocks), blkmap = {0, 1, 0}, nphys = 2, the operand from dfl_pack_conv_weights_split and
 * DFL_CONV_SPLIT_IO (outputs written as pairs, residual read as hi + lo); fp32 accumulation in TMEM.  Weight gradients
 * are three dfl_conv3x3_wgrad launches (hi.hi, lo.hi, hi.lo) accumulating into the same fp32 buffer. */
int dfl_conv3x3_fwd_ex(const void* x, const void* w_packed, const float* bias, void* out, void* out2,
                       const void* residual, const void* mask_src, const int64_t* dims, int ndim, int cin, int cout,
                       int flags, const int32_t* blkmap, int nphys, void* stream);
int dfl_pack_conv_weights_split(const float* w, void* w_fwd, void* w_dgrad, int taps, int cin, int cout, void* stream);
/* fp32 [n][cin] -> (hi, lo) bf16 [2][n][cpad] (channels >= cin zero);  (hi, lo) bf16 [2][n] -> fp32 [n] */
int dfl_split_f32(const float* in, void* out, size_t n, int cin, int cpad, void* stream);
int dfl_merge_split(const void* in, float* out, size_t n, void* stream);
/* dfl_pool_mask on (hi, lo) pairs (inputs and outputs) */
int dfl_pool_mask_split(const void* g, const void* mask_src, void* ds, void* dmasked, const int64_t* cdims, int ndim,
                        void* stream);

/* ---- optimizer (tf.train.AdamOptimizer / GradientDescentOptimizer: trainer.py:160-165) ----------------- */
/* flat fp32 buffers; lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed by the caller; m = v = NULL selects plain GD. */
int dfl_adam_step(float* param, const float* grad, float* m, float* v, size_t n, float lr_t, float beta1, float beta2,
                  float eps, float grad_scale, void* stream);

/* same, with the step size read from device memory (`lr_t_dev[0]`) so the launch can live inside a CUDA graph */
int dfl_adam_step_dev(float* param, const float* grad, float* m, float* v, size_t n, const float* lr_t_dev, float beta1,
                      float beta2, float eps, float grad_scale, void* stream);

/* ---- latent-space MLP of arch=nn (reference model.py:218-224, ops.py:26-36) ------------------------------------------
 * slim.batch_norm(decay, epsilon, scale=True, fused=True, is_training, activation_fn) on [M, N] fp32 + the activation:
 * training: normalise with the batch statistics (biased variance), update moving_mean / moving_var in place (the variance
 * Bessel-corrected, as fused batch norm returns it), keep save_mean / save_rstd for the backward; inference: moving
 * statistics.  act: 0 none, 1 leaky-ReLU 0.2 (ops.batch_norm's default), 2 ELU (NN's default tf.nn.elu).
 * dfl_bn_act_bwd: dx (may be NULL), dgamma, dbeta (overwritten) from dy = d loss / d y. */
int dfl_bn_act_fwd(const float* x, const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* y,
                   float* save_mean, float* save_rstd, int M, int N, float eps, float decay, int training, int act,
                   void* stream);
int dfl_bn_act_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* save_mean,
                   const float* save_rstd, float* dx, float* dgamma, float* dbeta, int M, int N, int act, void* stream);
/* slim.dropout(x, keep_prob, is_training=True): y = x * mask / keep_prob, mask from a counter-based generator keyed by
 * (seed, offset + index); the backward pass is the same call on dy with the same (seed, offset). */
int dfl_dropout(const float* x, float* y, size_t n, float keep_prob, uint64_t seed, uint64_t offset, void* stream);

/* ---- data-parallel exchange (new -- the reference is single-GPU) ------------------------------------
 * One process per GPU, every rank a full replica, ONE in-place ncclAllReduce(sum) over the flat fp32 gradient buffer per
 * optimizer step; the 1/world (and 1/grad_accum) factor is the grad_scale of dfl_adam_step.  NCCL is dlopen'ed
 * (libnccl.so.2, preferring the copy already mapped into the process).  dfl_allreduce is asynchronous on `stream` and may
 * be captured into a CUDA graph together with the kernels around it.
 *   id128: 128 bytes (ncclUniqueId) created on rank 0 by dfl_comm_unique_id and distributed out of band;
 *   dfl_comm_init binds the communicator to the CURRENT CUDA device (collective: every rank calls it). */
int dfl_comm_unique_id(void* id128);
int dfl_comm_init(void** comm, int nranks, const void* id128, int rank);
int dfl_allreduce(void* buf, size_t count, int dtype, void* comm, void* stream);
int dfl_comm_destroy(void* comm);

/* ---- misc ------------------------------------------------------------------------------------------- */
int dfl_cast_f32_bf16(const float* in, void* out, size_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPFLUIDS_B200_H_ */
