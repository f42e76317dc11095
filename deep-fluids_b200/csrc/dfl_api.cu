// deepfluids_b200 -- extern "C" entry points declared in include/deepfluids_b200.h
#include <stdarg.h>
#include <string.h>

#include "dfl_common.cuh"

namespace dfl {

static thread_local char g_err[512] = "";
static int g_num_sms = 148;
static bool g_inited = false;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return DFL_OK;
  set_last_error("CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return DFL_ERR_CUDA;
}

int num_sms() { return g_num_sms; }

int encode_tensor_map(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* gaddr, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw,
                      const uint32_t* elem_strides) {
  if (!g_encode) {
    set_last_error("dfl_init() was not called (TMA driver entry point unresolved)");
    return DFL_ERR_INIT;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; if (elem_strides) estr[i] = elem_strides[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = g_encode(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(gaddr), gd, gs, bx, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (CUresult %d; rank %d, dims %llu %llu %llu..., box %u %u %u...)",
                   static_cast<int>(r), rank, (unsigned long long)gd[0], (unsigned long long)gd[1],
                   (unsigned long long)(rank > 2 ? gd[2] : 0), bx[0], bx[1], rank > 2 ? bx[2] : 0);
    return DFL_ERR_CUDA;
  }
  return DFL_OK;
}

// implemented in the other translation units
size_t stencil_loss_workspace_bytes(int nd, const int64_t* dims);
int stencil_loss_fwdbwd(int nd, const int64_t* dims, const void* pot, int pot_channels, const void* x, void* dpot,
                        void* vel, float* loss3, void* workspace, float w1, float w2, float grad_scale, int dt_pot,
                        int dt_x, int dpot_channels, cudaStream_t st);
int fwd_stencils(int op, int nd, const int64_t* dims, const void* in, int in_cs, void* out0, void* out1, int dtype,
                 cudaStream_t st);
int bwd_stencils(int op, int nd, const int64_t* dims, const float* g0, const float* g1, float* out, int out_cs,
                 cudaStream_t st);
int mse_loss(const float* d, float target, float* loss, float* dd, size_t n, float scale, cudaStream_t st);
size_t lastconv_curl_loss_bwd_workspace_bytes();
int lastconv_curl_loss_bwd_2d(const void* s, const float* pot, const float* x, const float* w, const void* mask_src, void* ds,
                              void* ds_masked, float* dw, float* db, float* dpot, float* vel, float* loss3, void* workspace,
                              const int64_t* dims, float w1, float w2, float grad_scale, cudaStream_t st);
int lastconv_curl_loss_bwd(const void* s, const float* pot, const float* x, const float* w, const void* mask_src, void* ds,
                           void* ds_masked, float* dw, float* db, float* dpot, float* vel, float* loss3, void* workspace,
                           const int64_t* dims, float w1, float w2, float grad_scale, cudaStream_t st);
int pack_phase_weights(const float* W, void* wf, void* wd, int nd, int cin, int cout, cudaStream_t st);
int comm_unique_id(void* id128);
int comm_init(void** comm, int nranks, const void* id128, int rank);
int allreduce(void* buf, size_t count, int dtype, void* comm, cudaStream_t st);
int comm_destroy(void* comm);
size_t l1_loss_workspace_bytes();
int l1_loss(const float* a, const float* b, float* loss, float* dd, size_t n, float scale, int accumulate, void* ws,
            cudaStream_t st);
int gemm_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int transA, int transB,
             int accumulate, cudaStream_t st);
int colsum_f32(const float* x, float* out, int M, int N, cudaStream_t st);
int conv_tc_launch(const void* x, const void* w_packed, const float* bias, void* out, void* out2,
                   const void* residual, const void* mask_src, const int64_t* dims, int nd, int cin, int cout,
                   int flags, const int32_t* blkmap, int nphys, cudaStream_t st);
int wgrad_tc_launch(const void* x, const void* dpre, float* dw, float* db, const int64_t* x_dims, const int64_t* dims,
                    int nd, int in_stride, int pad, int dw_tap_stride, int dw_row_stride, cudaStream_t st, int split = 0,
                    int ksize = 3);
int bn_act_fwd(const float* x, const float* gamma, const float* beta, float* mmean, float* mvar, float* y, float* save_mean,
               float* save_rstd, int M, int N, float eps, float decay, int training, int act, cudaStream_t st);
int bn_act_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* save_mean,
               const float* save_rstd, float* dx, float* dgamma, float* dbeta, int M, int N, int act, cudaStream_t st);
int dropout(const float* x, float* y, size_t n, float keep, unsigned long long seed, unsigned long long offset, cudaStream_t st);
int gather_stride2(const void* fine, void* coarse, const int64_t* cdims, int nd, cudaStream_t st);
int phase_wgrad_fold(const float* t64, float* dw, int nd, int cin, int cout, cudaStream_t st);
size_t deterministic_workspace_bytes();
int set_deterministic(void* ws, size_t bytes);
int conv_tap_launch(const void* x, const void* w_packed, const float* bias, void* out, void* out2,
                    const void* residual, const void* mask_src, const int64_t* in_dims, const int64_t* tile_dims,
                    const int64_t* out_dims, int nd, int cin, int in_stride, int ntap, const int32_t* taps,
                    int out_stride, const int32_t* out_off, int64_t w_ld, int flags, cudaStream_t st);
int pad_cast(const float* in, void* out, size_t n, int cin, cudaStream_t st);
int add_mask(const void* a, const void* b, const void* y, void* out, size_t n, cudaStream_t st);
int enc_fc_fwd(const void* flat, const float* W, const float* bias, float* z, int B, int V, int nblk, int Z, cudaStream_t st);
int enc_fc_bwd(const void* flat, const float* W, const float* dz, float* dW, float* db, void* dflat, int B, int V,
               int nblk, int Z, cudaStream_t st);
int fc_dz(const void* dout, const float* W, float* dz, int B, int K, int N, int dout_dtype, int accumulate, cudaStream_t st);
int ae_loss_p(const float* z, const float* y, float* dz, float* loss_p, int B, int Z, int P, float scale, cudaStream_t st);
int ae_sigmoid(const float* zl, float* z, int n, cudaStream_t st);
int pack_split(const float* W, void* wf, void* wd, int taps, int cin, int cout, cudaStream_t st);
int split_f32(const float* in, void* out, size_t n, int cin, int cpad, cudaStream_t st);
int merge_split(const void* in, float* out, size_t n, cudaStream_t st);
int pool_mask_split(const void* g, const void* mask_src, void* ds, void* dmasked, const int64_t* cdims, int nd,
                    cudaStream_t st);
int ae_sparse_bwd(const float* z, float* dz, float* dzl, float* loss_kl, int B, int Z, int P, float rho, float w5,
                  cudaStream_t st);
int fc_fwd(const float* z, const float* W, const float* bias, void* out, int B, int K, int N, int out_dtype,
           cudaStream_t st);
int fc_bwd(const float* z, const void* dout, float* dW, float* db, int B, int K, int N, int dout_dtype,
           cudaStream_t st);
int lastconv_fwd_tc(const void* s, const float* w, const float* bias, float* out, const int64_t* dims, int nd, int cout,
                    cudaStream_t st);
int lastconv_bwd_tc(const void* s, const float* dout, const float* w, const void* mask_src, void* ds, void* ds_masked,
                    float* dw, float* db, const int64_t* dims, int nd, int cout, cudaStream_t st);
int pool_mask(const void* g, const void* mask_src, void* ds, void* dmasked, const int64_t* cdims, int nd,
              cudaStream_t st, const void* addend = nullptr);
int bias_grad(const void* d, float* db, size_t npos, cudaStream_t st);
int pack_conv_weights(const float* W, void* wf, void* wd, int taps, int cin, int cout, int cin_ld, cudaStream_t st);
int pack_conv_weights_multi(const void* ptrs, int n_layers, int taps, int cin, int cout, cudaStream_t st);
int adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr_t, const float* lr_t_dev, float b1,
              float b2, float eps, float grad_scale, cudaStream_t st);
int cast_f32_bf16(const float* a, void* o, size_t n, cudaStream_t st);
int upscale2(const void* in, void* out, const int64_t* cdims, int nd, int channels, int dtype, cudaStream_t st);
int pool2(const void* g, void* out, const int64_t* cdims, int nd, int channels, int dtype, cudaStream_t st);

}  // namespace dfl

using namespace dfl;
#define ST(s) static_cast<cudaStream_t>(s)

#pragma GCC visibility push(default)
extern "C" {

int dfl_version(void) { return 100; }
const char* dfl_last_error(void) { return g_err; }

int dfl_init(int device) {
  int ndev = 0;
  DFL_CUDA_OK(cudaGetDeviceCount(&ndev));
  DFL_REQUIRE(device >= 0 && device < ndev, "dfl_init: device %d out of range (have %d)", device, ndev);
  DFL_CUDA_OK(cudaSetDevice(device));
  DFL_CUDA_OK(cudaFree(0));
  cudaDeviceProp prop;
  DFL_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_last_error("dfl_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                   prop.minor);
    return DFL_ERR_UNSUPPORTED;
  }
  g_num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  DFL_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) {
    set_last_error("dfl_init: cuTensorMapEncodeTiled not available from the driver");
    return DFL_ERR_INIT;
  }
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  g_inited = true;
  return DFL_OK;
}

int dfl_curl_fwd(const void* pot, void* vel, const int64_t* dims, int ndim, int pot_channels, int dtype, void* stream) {
  return fwd_stencils(0, ndim, dims, pot, pot_channels, vel, nullptr, dtype, ST(stream));
}
int dfl_jacobian_fwd(const void* vel, void* jac, void* vort_or_curl, const int64_t* dims, int ndim, int dtype,
                     void* stream) {
  return fwd_stencils(1, ndim, dims, vel, ndim, jac, vort_or_curl, dtype, ST(stream));
}
int dfl_divergence(const void* vel, void* div, const int64_t* dims, int ndim, int dtype, void* stream) {
  return fwd_stencils(2, ndim, dims, vel, ndim, div, nullptr, dtype, ST(stream));
}
int dfl_curl_bwd(const float* dvel, float* dpot, const int64_t* dims, int ndim, int dpot_channels, void* stream) {
  return bwd_stencils(0, ndim, dims, dvel, nullptr, dpot, dpot_channels, ST(stream));
}
int dfl_jacobian_bwd(const float* djac, const float* daux, float* dvel, const int64_t* dims, int ndim, void* stream) {
  return bwd_stencils(1, ndim, dims, djac, daux, dvel, ndim, ST(stream));
}
int dfl_mse_loss(const float* d, float target, float* loss, float* dd, size_t n, float scale, void* stream) {
  return mse_loss(d, target, loss, dd, n, scale, ST(stream));
}
size_t dfl_lastconv_curl_loss_workspace_bytes(void) { return lastconv_curl_loss_bwd_workspace_bytes(); }
int dfl_lastconv_curl_loss_fwd(const void* s, const float* w, const float* bias, float* pot, const int64_t* dims, int ndim,
                               int cout, void* stream) {
  return lastconv_fwd_tc(s, w, bias, pot, dims, ndim, cout, ST(stream));
}
int dfl_lastconv_curl_loss_bwd(const void* s, const float* pot, const float* x, const float* w, const void* mask_src,
                               void* ds, void* ds_masked, float* dw, float* db, float* dpot, float* vel, float* loss3,
                               void* workspace, const int64_t* dims, int ndim, float w1, float w2, float grad_scale,
                               void* stream) {
  DFL_REQUIRE(ndim == 2 || ndim == 3, "dfl_lastconv_curl_loss_bwd: ndim must be 2 or 3");
  if (ndim == 2)
    return lastconv_curl_loss_bwd_2d(s, pot, x, w, mask_src, ds, ds_masked, dw, db, dpot, vel, loss3, workspace, dims, w1, w2,
                                     grad_scale, ST(stream));
  return lastconv_curl_loss_bwd(s, pot, x, w, mask_src, ds, ds_masked, dw, db, dpot, vel, loss3, workspace, dims, w1, w2,
                                grad_scale, ST(stream));
}
int dfl_pack_phase_weights(const float* w, void* w_fwd, void* w_dgrad, int ndim, int cin, int cout, void* stream) {
  return pack_phase_weights(w, w_fwd, w_dgrad, ndim, cin, cout, ST(stream));
}
int dfl_comm_unique_id(void* id128) { return comm_unique_id(id128); }
int dfl_comm_init(void** comm, int nranks, const void* id128, int rank) { return comm_init(comm, nranks, id128, rank); }
int dfl_allreduce(void* buf, size_t count, int dtype, void* comm, void* stream) {
  return allreduce(buf, count, dtype, comm, ST(stream));
}
int dfl_comm_destroy(void* comm) { return comm_destroy(comm); }
size_t dfl_l1_loss_workspace_bytes(void) { return l1_loss_workspace_bytes(); }
int dfl_l1_loss(const float* a, const float* b, float* loss, float* dd, size_t n, float scale, int accumulate,
                void* workspace, void* stream) {
  return l1_loss(a, b, loss, dd, n, scale, accumulate, workspace, ST(stream));
}
int dfl_gemm_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int transA, int transB,
                 int accumulate, void* stream) {
  return gemm_f32(A, B, bias, C, M, N, K, transA, transB, accumulate, ST(stream));
}
int dfl_colsum_f32(const float* x, float* out, int M, int N, void* stream) { return colsum_f32(x, out, M, N, ST(stream)); }
size_t dfl_stencil_loss_workspace_bytes(const int64_t* dims, int ndim) {
  return stencil_loss_workspace_bytes(ndim, dims);
}
int dfl_stencil_loss_fwdbwd(const void* pot, const void* x, void* dpot, void* vel, float* loss3, void* workspace,
                            const int64_t* dims, int ndim, int pot_channels, float w1, float w2, float grad_scale,
                            int dtype_pot, int dtype_x, void* stream) {
  return stencil_loss_fwdbwd(ndim, dims, pot, pot_channels, x, dpot, vel, loss3, workspace, w1, w2, grad_scale,
                             dtype_pot, dtype_x, 0, ST(stream));
}
int dfl_stencil_loss_fwdbwd_ex(const void* pot, const void* x, void* dpot, void* vel, float* loss3, void* workspace,
                               const int64_t* dims, int ndim, int pot_channels, int dpot_channels, float w1, float w2,
                               float grad_scale, int dtype_pot, int dtype_x, void* stream) {
  return stencil_loss_fwdbwd(ndim, dims, pot, pot_channels, x, dpot, vel, loss3, workspace, w1, w2, grad_scale,
                             dtype_pot, dtype_x, dpot_channels, ST(stream));
}
int dfl_fc_fwd(const float* z, const float* W, const float* bias, void* out, int B, int K, int N, int out_dtype,
               void* stream) {
  return fc_fwd(z, W, bias, out, B, K, N, out_dtype, ST(stream));
}
int dfl_fc_bwd(const float* z, const void* dout, float* dW, float* db, int B, int K, int N, int dout_dtype,
               void* stream) {
  return fc_bwd(z, dout, dW, db, B, K, N, dout_dtype, ST(stream));
}
int dfl_pack_conv_weights(const float* w, void* w_fwd, void* w_dgrad, int taps, int cin, int cout, void* stream) {
  return pack_conv_weights(w, w_fwd, w_dgrad, taps, cin, cout, cin, ST(stream));
}
int dfl_pack_conv_weights_multi(const void* ptr_table, int n_layers, int taps, int cin, int cout, void* stream) {
  return pack_conv_weights_multi(ptr_table, n_layers, taps, cin, cout, ST(stream));
}
int dfl_pack_conv_weights_ex(const float* w, void* w_fwd, void* w_dgrad, int taps, int cin, int cout, int cin_ld,
                             void* stream) {
  return pack_conv_weights(w, w_fwd, w_dgrad, taps, cin, cout, cin_ld, ST(stream));
}
int dfl_conv3x3_fwd(const void* x, const void* w_packed, const float* bias, void* out, void* out2,
                    const void* residual, const void* mask_src, const int64_t* dims, int ndim, int cin, int cout,
                    int flags, void* stream) {
  return conv_tc_launch(x, w_packed, bias, out, out2, residual, mask_src, dims, ndim, cin, cout, flags, nullptr, 0,
                        ST(stream));
}
int dfl_conv3x3_fwd_ex(const void* x, const void* w_packed, const float* bias, void* out, void* out2,
                       const void* residual, const void* mask_src, const int64_t* dims, int ndim, int cin, int cout,
                       int flags, const int32_t* blkmap, int nphys, void* stream) {
  return conv_tc_launch(x, w_packed, bias, out, out2, residual, mask_src, dims, ndim, cin, cout, flags, blkmap, nphys,
                        ST(stream));
}
int dfl_conv3x3_wgrad(const void* x, const void* dpre, float* dw, float* db, const int64_t* dims, int ndim, int cin,
                      int cout, void* stream) {
  DFL_REQUIRE(cin == 128 && cout == 128, "conv3x3_wgrad: Cin = Cout = 128 only (use dfl_conv_wgrad_ex for blocks)");
  return wgrad_tc_launch(x, dpre, dw, db, dims, dims, ndim, 1, 1, 128 * 128, 128, ST(stream));
}
int dfl_conv3x3_wgrad_split(const void* x2, const void* dpre2, float* dw, float* db, const int64_t* dims, int ndim,
                            void* stream) {
  return wgrad_tc_launch(x2, dpre2, dw, db, dims, dims, ndim, 1, 1, 128 * 128, 128, ST(stream), 1);
}
int dfl_conv_wgrad_ex(const void* x, const void* dpre, float* dw, float* db, const int64_t* x_dims, const int64_t* dims,
                      int ndim, int in_stride, int pad, int dw_tap_stride, int dw_row_stride, void* stream) {
  return wgrad_tc_launch(x, dpre, dw, db, x_dims, dims, ndim, in_stride, pad, dw_tap_stride, dw_row_stride, ST(stream));
}
int dfl_phase_wgrad(const void* dy_fine, const void* s_coarse, float* t_scratch, const int64_t* fine_dims,
                    const int64_t* coarse_dims, int ndim, void* stream) {
  return wgrad_tc_launch(dy_fine, s_coarse, t_scratch, nullptr, fine_dims, coarse_dims, ndim, 2, 1, 128 * 128, 128, ST(stream),
                         0, 4);
}
int dfl_bn_act_fwd(const float* x, const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* y,
                   float* save_mean, float* save_rstd, int M, int N, float eps, float decay, int training, int act,
                   void* stream) {
  return bn_act_fwd(x, gamma, beta, moving_mean, moving_var, y, save_mean, save_rstd, M, N, eps, decay, training, act, ST(stream));
}
int dfl_bn_act_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* save_mean,
                   const float* save_rstd, float* dx, float* dgamma, float* dbeta, int M, int N, int act, void* stream) {
  return bn_act_bwd(x, y, dy, gamma, save_mean, save_rstd, dx, dgamma, dbeta, M, N, act, ST(stream));
}
int dfl_dropout(const float* x, float* y, size_t n, float keep_prob, uint64_t seed, uint64_t offset, void* stream) {
  return dropout(x, y, n, keep_prob, seed, offset, ST(stream));
}
int dfl_gather_stride2(const void* fine, void* coarse, const int64_t* cdims, int ndim, void* stream) {
  return gather_stride2(fine, coarse, cdims, ndim, ST(stream));
}
size_t dfl_deterministic_workspace_bytes(void) { return deterministic_workspace_bytes(); }
int dfl_set_deterministic(void* workspace, size_t bytes) { return set_deterministic(workspace, bytes); }
int dfl_phase_wgrad_fold(const float* t_scratch, float* dw, int ndim, int cin, int cout, void* stream) {
  return phase_wgrad_fold(t_scratch, dw, ndim, cin, cout, ST(stream));
}
int dfl_conv_taps(const void* x, const void* w_packed, const float* bias, void* out, void* out2, const void* residual,
                  const void* mask_src, const int64_t* in_dims, const int64_t* tile_dims, const int64_t* out_dims, int ndim,
                  int cin, int in_stride, int ntap, const int32_t* taps, int out_stride, const int32_t* out_off,
                  int64_t w_ld, int flags, void* stream) {
  return conv_tap_launch(x, w_packed, bias, out, out2, residual, mask_src, in_dims, tile_dims, out_dims, ndim, cin,
                         in_stride, ntap, taps, out_stride, out_off, w_ld, flags, ST(stream));
}
int dfl_pad_cast(const float* in, void* out, size_t n, int cin, void* stream) { return pad_cast(in, out, n, cin, ST(stream)); }
int dfl_add_mask(const void* a, const void* b, const void* y, void* out, size_t n, void* stream) {
  return add_mask(a, b, y, out, n, ST(stream));
}
int dfl_enc_fc_fwd(const void* flat, const float* W, const float* bias, float* z, int B, int V, int nblk, int Z,
                   void* stream) {
  return enc_fc_fwd(flat, W, bias, z, B, V, nblk, Z, ST(stream));
}
int dfl_enc_fc_bwd(const void* flat, const float* W, const float* dz, float* dW, float* db, void* dflat, int B, int V,
                   int nblk, int Z, void* stream) {
  return enc_fc_bwd(flat, W, dz, dW, db, dflat, B, V, nblk, Z, ST(stream));
}
int dfl_fc_dz(const void* dout, const float* W, float* dz, int B, int K, int N, int dout_dtype, int accumulate,
              void* stream) {
  return fc_dz(dout, W, dz, B, K, N, dout_dtype, accumulate, ST(stream));
}
int dfl_ae_loss_p(const float* z, const float* y, float* dz, float* loss_p, int B, int Z, int P, float scale,
                  void* stream) {
  return ae_loss_p(z, y, dz, loss_p, B, Z, P, scale, ST(stream));
}
int dfl_bias_grad(const void* dpre, float* db, size_t npos, void* stream) {
  return bias_grad(dpre, db, npos, ST(stream));
}
int dfl_lastconv_fwd(const void* s, const float* w, const float* bias, float* out, const int64_t* dims, int ndim, int cout,
                     void* stream) {
  return lastconv_fwd_tc(s, w, bias, out, dims, ndim, cout, ST(stream));
}
int dfl_lastconv_bwd(const void* s, const float* dout, const float* w, const void* mask_src, void* ds, void* ds_masked,
                     float* dw, float* db, const int64_t* dims, int ndim, int cout, void* stream) {
  return lastconv_bwd_tc(s, dout, w, mask_src, ds, ds_masked, dw, db, dims, ndim, cout, ST(stream));
}
int dfl_pool_mask(const void* g, const void* mask_src, void* ds, void* dmasked, const int64_t* cdims, int ndim,
                  void* stream) {
  DFL_REQUIRE(!(dmasked && !mask_src), "pool_mask: dmasked requested without mask_src");
  return pool_mask(g, mask_src, ds, dmasked, cdims, ndim, ST(stream));
}
int dfl_pool_mask_add(const void* g, const void* addend, const void* mask_src, void* ds, void* dmasked, const int64_t* cdims,
                      int ndim, void* stream) {
  return pool_mask(g, mask_src, ds, dmasked, cdims, ndim, ST(stream), addend);
}
int dfl_adam_step(float* param, const float* grad, float* m, float* v, size_t n, float lr_t, float beta1, float beta2,
                  float eps, float grad_scale, void* stream) {
  return adam_step(param, grad, m, v, n, lr_t, nullptr, beta1, beta2, eps, grad_scale, ST(stream));
}
int dfl_adam_step_dev(float* param, const float* grad, float* m, float* v, size_t n, const float* lr_t_dev, float beta1,
                      float beta2, float eps, float grad_scale, void* stream) {
  DFL_REQUIRE(lr_t_dev != nullptr, "adam_step_dev: lr_t_dev is NULL");
  return adam_step(param, grad, m, v, n, 0.f, lr_t_dev, beta1, beta2, eps, grad_scale, ST(stream));
}
int dfl_upscale2(const void* in, void* out, const int64_t* cdims, int ndim, int channels, int dtype, void* stream) {
  return upscale2(in, out, cdims, ndim, channels, dtype, ST(stream));
}
int dfl_pool2(const void* g, void* out, const int64_t* cdims, int ndim, int channels, int dtype, void* stream) {
  return pool2(g, out, cdims, ndim, channels, dtype, ST(stream));
}
int dfl_cast_f32_bf16(const float* in, void* out, size_t n, void* stream) {
  return cast_f32_bf16(in, out, n, ST(stream));
}

int dfl_pack_conv_weights_split(const float* w, void* w_fwd, void* w_dgrad, int taps, int cin, int cout, void* stream) {
  return pack_split(w, w_fwd, w_dgrad, taps, cin, cout, ST(stream));
}
int dfl_split_f32(const float* in, void* out, size_t n, int cin, int cpad, void* stream) {
  return split_f32(in, out, n, cin, cpad, ST(stream));
}
int dfl_merge_split(const void* in, float* out, size_t n, void* stream) { return merge_split(in, out, n, ST(stream)); }
int dfl_pool_mask_split(const void* g, const void* mask_src, void* ds, void* dmasked, const int64_t* cdims, int ndim,
                        void* stream) {
  DFL_REQUIRE(!(dmasked && !mask_src), "pool_mask_split: dmasked requested without mask_src");
  return pool_mask_split(g, mask_src, ds, dmasked, cdims, ndim, ST(stream));
}
int dfl_ae_sigmoid(const float* zl, float* z, int n, void* stream) { return ae_sigmoid(zl, z, n, ST(stream)); }
int dfl_ae_sparse_bwd(const float* z, float* dz, float* dzl, float* loss_kl, int B, int Z, int P, float rho, float w5,
                      void* stream) {
  return ae_sparse_bwd(z, dz, dzl, loss_kl, B, Z, P, rho, w5, ST(stream));
}

}  // extern "C"
#pragma GCC visibility pop
