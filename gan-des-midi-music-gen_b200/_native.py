"""ctypes binding of libmmgan_b200.so (C ABI declared in include/mmgan_b200.h).

There is no CPU fallback: if the library is missing, or a tensor is not on a CUDA device, the
call raises.  PyTorch is only used for device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmgan_b200.so")

_P, _I, _L, _F, _Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); must list every function of include/mmgan_b200.h (tests/test_abi.py checks)
SIGNATURES = {
    "mmg_abi_version": (_I, []),
    "mmg_last_error": (ctypes.c_char_p, []),
    "mmg_launch_count": (ctypes.c_uint64, []),
    "mmg_raster_out_width": (_I, [_I, _I]),
    "mmg_raster_workspace_bytes": (_Z, [_L, _L]),
    "mmg_raster_set_mode": (_I, [_I]),
    "mmg_smf_max_messages": (_L, [_Z]),
    "mmg_smf_parse": (_I, [_P, _Z, _P, _P, _P, _L, _P, _P, _P, _P, _L, _P]),
    "mmg_smf_beat_grid": (_I, [_P, _P, _L, _I, _L, _P, _L, _P]),
    "mmg_simlog_max_messages": (_I, []),
    "mmg_simlog_to_events": (_I, [_P, _Z, _P, _I, _P, _I, _P, _I, _I, _I, _P, _P, _L, _P]),
    "mmg_simlog_batch_to_events": (_I, [_P, _P, _L, _P, _I, _P, _I, _P, _I, _I, _I, _P, _P, _P, _I]),
    "mmg_raster_piano_roll": (_I, [_P, _P, _P, _L, _L, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "mmg_bce_logits_f32": (_I, [_P, _P, _F, _L, _P, _I, _P, _F, _P, _P]),
    "mmg_fill_scalar_f32": (_I, [_P, _P, _L, _P]),
    "mmg_zero": (_I, [_P, _Z, _P]),
    "mmg_sum_f32": (_I, [_P, _L, _P, _I, _P]),
    "mmg_act_bwd_f32": (_I, [_P, _P, _P, _L, _I, _P]),
    "mmg_adam_multi_tensor_f32": (_I, [_I, _P, _P, _F, _F, _F, _F, _L, _F, _P]),
    "mmg_adam_multi_tensor_dev_f32": (_I, [_I, _P, _P, _P, _P, _F, _P]),
    "mmg_linear_fwd_f32": (_I, [_P, _P, _P, _P, _L, _L, _L, _I, _P]),
    "mmg_linear_bwd_f32": (_I, [_P, _P, _P, _P, _P, _P, _L, _L, _L, _I, _P]),
    "mmg_bn_workspace_bytes": (_Z, [_L]),
    "mmg_bn_fwd_train_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _L, _F, _F, _I, _P, _Z, _P]),
    "mmg_bn_fwd_eval_f32": (_I, [_P, _P, _P, _P, _P, _P, _L, _L, _L, _F, _I, _P]),
    "mmg_bn_bwd_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _L, _I, _I, _P, _Z, _P]),
    "mmg_conv2d_fwd_f32": (_I, [_P, _P, _P, _P] + [_I] * 10 + [_P]),
    "mmg_conv2d_bwd_data_f32": (_I, [_P, _P, _P, _P] + [_I] * 10 + [_P]),
    "mmg_conv2d_bwd_weight_f32": (_I, [_P, _P, _P, _P] + [_I] * 10 + [_P]),
    "mmg_gemm_tc": (_I, [_P, _I, _L, _P, _I, _L, _P, _L, _I, _I, _I, _I, _I, _I, _L, _I, _P, _I, _I, _P]),
    "mmg_pack_bf16": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _L, _L, _P]),
    "mmg_im2col_bf16": (_I, [_P, _P] + [_I] * 10 + [_P]),
    "mmg_col2im_f32": (_I, [_P, _P] + [_I] * 8 + [_L, _I, _P]),
    "mmg_transpose_f32": (_I, [_P, _P, _I, _I, _L, _P]),
    "mmg_colsum_f32": (_I, [_P, _P, _I, _I, _P]),
    "mmg_bias_act_inplace_f32": (_I, [_P, _P, _L, _I, _I, _P]),
    "mmg_conv_small_relu_pool_f32": (_I, [_P, _P, _P, _P, _P] + [_I] * 8 + [_P]),
    "mmg_pool_relu_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _L, _P]),
    "mmg_conv_dgrad_gather": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _L, _P]),
    "mmg_stft_power_f32": (_I, [_P, _I, _L, _L, _I, _I, _P, _I, _P]),
    "mmg_power_to_db_f32": (_I, [_P, _P, _I, _L, _F, _P]),
    "mmg_maxpool2_fwd_f32": (_I, [_P, _P, _P, _L, _I, _I, _P]),
    "mmg_maxpool2_bwd_f32": (_I, [_P, _P, _P, _L, _I, _I, _P]),
    "mmg_disc_packed_weights_bytes": (_Z, []),
    "mmg_disc_pack_weights": (_I, [_P, _P, _P, _P, _P]),
    "mmg_disc_xs_pack": (_I, [_P, _I, _P, _L, _P]),
    "mmg_disc_conv1_fwd": (_I, [_P, _P, _P, _P, _L, _P]),
    "mmg_disc_conv2_fwd": (_I, [_P, _P, _P, _P, _P, _L, _P]),
    "mmg_disc_fc_bwd": (_I, [_P, _P, _P, _P, _P, _P, _L, _P]),
    "mmg_disc_conv2_wgrad": (_I, [_P, _P, _P, _L, _P]),
    "mmg_disc_conv2_dgrad": (_I, [_P, _P, _P, _P, _P, _L, _P]),
    "mmg_disc_conv1_wgrad": (_I, [_P, _P, _P, _L, _P]),
}



class GenLayerArgs(ctypes.Structure):
    """mmg_gen_layer_args of include/mmgan_b200.h"""
    _fields_ = [("x0", _P), ("x1", _P), ("k0", _I), ("k1", _I),
                ("in_mode", _I), ("in_sums", _P), ("in_gamma", _P), ("in_beta", _P), ("in_run_mean", _P), ("in_run_var", _P),
                ("w_packed", _P), ("bias", _P), ("N", _I),
                ("z_out", _P), ("out_sums", _P), ("y_out", _P),
                ("out_mode", _I), ("y_sums", _P), ("out_gamma", _P), ("out_beta", _P), ("out_run_mean", _P), ("out_run_var", _P),
                ("momentum", _F), ("eps", _F), ("update_running", _I), ("M", _L), ("stat_count", _L)]


class GenHiddenArgs(ctypes.Structure):
    """mmg_gen_hidden_args of include/mmgan_b200.h"""
    _fields_ = [("x0", _P), ("x1", _P), ("k0", _I), ("k1", _I),
                ("w_packed", _P * 3), ("bias", _P * 3), ("N", _I * 3),
                ("gamma", _P * 3), ("beta", _P * 3), ("run_mean", _P * 3), ("run_var", _P * 3),
                ("sums", _P * 3),
                ("z_out", _P), ("gram_part", _P), ("barrier", _P),
                ("momentum", _F), ("eps", _F), ("update_running", _I), ("M", _L), ("stat_count", _L)]


SIGNATURES.update({
    "mmg_disc_fwd_fused": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P]),
    "mmg_disc_fwd_fused_gather": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P]),
    "mmg_disc_bwd_fused": (_I, [_P] * 11 + [_L, _P]),
    "mmg_disc_pass_workspace_bytes": (_Z, []),
    "mmg_disc_pass_set_flags": (_I, [_I]),
    "mmg_disc_pass_fused": (_I, [_P, _I, _P, _L, _P, _P, _P, _P, _F, _L, _P, _P] + [_P] * 6 + [_P, _Z, _L, _P]),
    "mmg_disc_pass_fused_dbg": (_I, [_P, _I, _P, _L, _P, _P, _P, _P, _F, _L, _P, _P] + [_P] * 6 + [_P, _Z, _L, _P, _P]),
    "mmg_gen_packed_weight_bytes": (_Z, [_I, _I]),
    "mmg_gen_pack_weight": (_I, [_P, _I, _I, _P, _P]),
    "mmg_gen_layer_fwd": (_I, [ctypes.POINTER(GenLayerArgs), _P]),
    "mmg_gen_set_worker_groups": (_I, [_I]),
    "mmg_gen_hidden_fused_supported": (_I, [_L, _I, ctypes.POINTER(_I), _I]),
    "mmg_gen_hidden_fused": (_I, [ctypes.POINTER(GenHiddenArgs), _P]),
    "mmg_gen_layer_stats_gram_finish": (_I, [_I, _P, _P, _I, _I, _L, _P, _P, _Z, _P]),
    "mmg_gen_layer_stats_gram_workspace": (_Z, []),
    "mmg_gen_layer_stats_gram": (_I, [_P, _L, _I, _P, _L, _P, _P, _F, _P, _P, _I, _P, _P, _Z, _P]),
})

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Loads the library once; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first; there is no CPU fallback")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().mmg_last_error().decode(errors="replace")
        if rc == -1 and "Expected more than 1 value per channel" in msg:
            raise ValueError(msg)
        if rc < 0:
            raise ValueError(f"{what}: {msg} (code {rc})")
        raise NativeError(f"{what}: {msg} (cudaError {rc})")


def call(name, *args):
    check(getattr(lib(), name)(*args), name)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("mmgan_b200 kernels take CUDA tensors only (there is no CPU fallback)")
    if not t.is_contiguous():
        raise NativeError("mmgan_b200 kernels take contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NativeError("mmgan_b200 modules run on CUDA tensors only (there is no CPU fallback); got a CPU tensor")
