"""ctypes binding of include/deepfluids_b200.h (the drop-in boundary).  No torch types cross this line: only raw
device pointers, sizes and a cudaStream_t.  Fails loudly when the library is missing -- there is no fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFL_LIB_PATH") or os.path.join(_HERE, "lib", "libdeepfluids_b200.so")   # override: kernel A/B experiments

F32, BF16 = 0, 1
CONV_LRELU, CONV_OUT2_UPSAMPLE, CONV_MASK_AFTER_RESIDUAL, CONV_SPLIT_IO = 1, 2, 4, 8

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_dims = C.POINTER(C.c_int64)

# name -> (restype, argtypes): one row per declaration in include/deepfluids_b200.h
SIGNATURES = {
    "dfl_version": (_i, []),
    "dfl_last_error": (C.c_char_p, []),
    "dfl_init": (_i, [_i]),
    "dfl_curl_fwd": (_i, [_vp, _vp, _dims, _i, _i, _i, _vp]),
    "dfl_jacobian_fwd": (_i, [_vp, _vp, _vp, _dims, _i, _i, _vp]),
    "dfl_divergence": (_i, [_vp, _vp, _dims, _i, _i, _vp]),
    "dfl_curl_bwd": (_i, [_vp, _vp, _dims, _i, _i, _vp]),
    "dfl_jacobian_bwd": (_i, [_vp, _vp, _vp, _dims, _i, _vp]),
    "dfl_mse_loss": (_i, [_vp, _f, _vp, _vp, _sz, _f, _vp]),
    "dfl_lastconv_curl_loss_workspace_bytes": (_sz, []),
    "dfl_lastconv_curl_loss_fwd": (_i, [_vp, _vp, _vp, _vp, _dims, _i, _i, _vp]),
    "dfl_lastconv_curl_loss_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _dims, _i, _f, _f, _f, _vp]),
    "dfl_comm_unique_id": (_i, [_vp]),
    "dfl_comm_init": (_i, [C.POINTER(_vp), _i, _vp, _i]),
    "dfl_allreduce": (_i, [_vp, _sz, _i, _vp, _vp]),
    "dfl_comm_destroy": (_i, [_vp]),
    "dfl_l1_loss_workspace_bytes": (_sz, []),
    "dfl_l1_loss": (_i, [_vp, _vp, _vp, _vp, _sz, _f, _i, _vp, _vp]),
    "dfl_gemm_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "dfl_colsum_f32": (_i, [_vp, _vp, _i, _i, _vp]),
    "dfl_stencil_loss_workspace_bytes": (_sz, [_dims, _i]),
    "dfl_stencil_loss_fwdbwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _dims, _i, _i, _f, _f, _f, _i, _i, _vp]),
    "dfl_stencil_loss_fwdbwd_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _dims, _i, _i, _i, _f, _f, _f, _i, _i, _vp]),
    "dfl_fc_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dfl_fc_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dfl_pack_conv_weights": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "dfl_pack_conv_weights_multi": (_i, [_vp, _i, _i, _i, _i, _vp]),
    "dfl_conv3x3_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _dims, _i, _i, _i, _i, _vp]),
    "dfl_conv3x3_wgrad": (_i, [_vp, _vp, _vp, _vp, _dims, _i, _i, _i, _vp]),
    "dfl_conv3x3_wgrad_split": (_i, [_vp, _vp, _vp, _vp, _dims, _i, _vp]),
    "dfl_bias_grad": (_i, [_vp, _vp, _sz, _vp]),
    "dfl_lastconv_fwd": (_i, [_vp, _vp, _vp, _vp, _dims, _i, _i, _vp]),
    "dfl_lastconv_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _dims, _i, _i, _vp]),
    "dfl_pack_conv_weights_ex": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dfl_conv_taps": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _dims, _dims, _dims, _i, _i, _i, _i, C.POINTER(C.c_int32), _i,
                           C.POINTER(C.c_int32), C.c_int64, _i, _vp]),
    "dfl_conv_wgrad_ex": (_i, [_vp, _vp, _vp, _vp, _dims, _dims, _i, _i, _i, _i, _i, _vp]),
    "dfl_pad_cast": (_i, [_vp, _vp, _sz, _i, _vp]),
    "dfl_add_mask": (_i, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "dfl_enc_fc_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dfl_enc_fc_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dfl_fc_dz": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "dfl_ae_loss_p": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "dfl_ae_sigmoid": (_i, [_vp, _vp, _i, _vp]),
    "dfl_ae_sparse_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _vp]),
    "dfl_conv3x3_fwd_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _dims, _i, _i, _i, _i, C.POINTER(C.c_int32), _i, _vp]),
    "dfl_pack_conv_weights_split": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "dfl_split_f32": (_i, [_vp, _vp, _sz, _i, _i, _vp]),
    "dfl_merge_split": (_i, [_vp, _vp, _sz, _vp]),
    "dfl_pool_mask_split": (_i, [_vp, _vp, _vp, _vp, _dims, _i, _vp]),
    "dfl_pool_mask": (_i, [_vp, _vp, _vp, _vp, _dims, _i, _vp]),
    "dfl_pack_phase_weights": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "dfl_phase_wgrad": (_i, [_vp, _vp, _vp, _dims, _dims, _i, _vp]),
    "dfl_bn_act_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _i, _i, _vp]),
    "dfl_bn_act_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "dfl_dropout": (_i, [_vp, _vp, _sz, _f, C.c_uint64, C.c_uint64, _vp]),
    "dfl_gather_stride2": (_i, [_vp, _vp, _dims, _i, _vp]),
    "dfl_phase_wgrad_fold": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "dfl_upscale2": (_i, [_vp, _vp, _dims, _i, _i, _i, _vp]),
    "dfl_pool2": (_i, [_vp, _vp, _dims, _i, _i, _i, _vp]),
    "dfl_deterministic_workspace_bytes": (C.c_size_t, []),
    "dfl_set_deterministic": (_i, [_vp, C.c_size_t]),
    "dfl_pool_mask_add": (_i, [_vp, _vp, _vp, _vp, _vp, _dims, _i, _vp]),
    "dfl_adam_step": (_i, [_vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _vp]),
    "dfl_adam_step_dev": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _f, _f, _f, _f, _vp]),
    "dfl_cast_f32_bf16": (_i, [_vp, _vp, _sz, _vp]),
}


class DflError(RuntimeError):
    pass


_lib = None
_inited_device = None


def load():
    """dlopen the library and attach prototypes (no CUDA call is made)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DflError("deepfluids_b200: %s not found -- build it with `python __graft_entry__.py build` "
                           "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def lib(device=None):
    """Loaded library, initialised for `device` (default: torch's current CUDA device)."""
    global _inited_device
    l = load()
    if device is None:
        import torch
        if not torch.cuda.is_available():
            raise DflError("deepfluids_b200: no CUDA device visible; the hot path is B200-only (no CPU fallback)")
        device = torch.cuda.current_device()
    if _inited_device != device:
        check(l.dfl_init(int(device)))
        _inited_device = device
    return l


def check(rc):
    if rc != 0:
        raise DflError("deepfluids_b200 error %d: %s" % (rc, load().dfl_last_error().decode()))


def dims_array(shape):
    return (C.c_int64 * len(shape))(*[int(s) for s in shape])
