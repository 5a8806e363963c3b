"""ctypes binding of libsedb200.so (the C ABI in include/sed_b200.h).

ctypes releases the GIL for the duration of every call, so DataParallel's per-GPU Python
threads enqueue concurrently.  All functions return an int status; ``call`` turns a non-zero
status into ``RuntimeError(sed_last_error_string())``.
"""
import ctypes
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# SED_B200_LIB: development override (A/B builds made by build.build_variant); the product loads the in-tree library
LIB_PATH = os.environ.get('SED_B200_LIB') or os.path.join(_PKG_DIR, 'libsedb200.so')

_c_int, _c_ll, _c_float, _c_void_p = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p
P, I, L, F, D = _c_void_p, _c_int, _c_ll, _c_float, ctypes.c_double
U = ctypes.c_ulonglong

# name -> argtypes (all return int unless listed in _RESTYPES)
SIGNATURES = {
    'sed_abi_version': [],
    'sed_launch_count': [],
    'sed_last_error_string': [],
    'sed_device_sm_count': [P],
    'sed_logmel_f32': [P, I, I, I, P, P, P, I, F, F, P, P],
    'sed_logmel_i16': [P, I, I, I, P, P, P, I, F, F, P, P],
    'sed_stft_power_f32': [P, I, I, I, P, P],
    'sed_mel_db_f32': [P, L, I, P, P, P, I, F, F, I, P, P],
    'sed_conv_pack_weights': [P, I, I, P, P, P],
    'sed_conv3x3_tc_grid': [I, I, I, I, I],
    'sed_conv3x3_tc_fwd': [P, P, P, P, I, I, I, I, I, P],
    'sed_conv3x3_tc2_grid': [I, I, I, I, I],
    'sed_conv3x3_tc2_fwd': [P, P, P, P, I, I, I, I, I, P],
    'sed_conv3x3_tc2kw_supported': [I, I, I],
    'sed_conv3x3_tc2kw_fwd': [P, P, P, P, I, I, I, I, I, P],
    'sed_conv3x3_tc_dgrad_bnr': [P, P, P, I, I, I, I, I, P, I, P, P, I, P, P],
    'sed_conv3x3_tc_wgrad_splits': [I, I, I, I, I],
    'sed_conv3x3_tc_wgrad_use_pairs': [I],
    'sed_conv3x3_tc_wgrad': [P, P, P, I, I, I, I, I, P],
    'sed_conv_unpack_wgrad': [P, I, L, I, I, P, I, P],
    'sed_conv_pack_weights_multi': [I, P, P, P, P, P, P],
    'sed_conv_unpack_wgrad_multi': [I, P, P, P, P, P, P],
    'sed_f32_to_bf16': [P, P, L, P],
    'sed_bn_finalize': [P, I, I, D, P, P, F, F, P, P, P, P, P, P, P, P],
    'sed_bn_eval_affine': [P, P, P, P, F, I, P, P, P, P, P],
    'sed_bn_relu_pool_fwd': [P, P, P, I, I, I, I, I, I, P, I, P],
    'sed_bn_bwd_partials': [I, I, I, I, I, I],
    'sed_bn_bwd_workspace_rows': [I, I, I, I, I, I],
    'sed_bn_bwd_sched_words': [],
    'sed_bn_relu_pool_bwd_reduce': [P, P, I, P, P, I, I, I, I, I, I, P, P, P],
    'sed_bn_bwd_finalize': [P, I, I, D, P, P, P, P, P, I, I, P, P],
    'sed_bn_relu_pool_bwd_apply': [P, P, I, P, P, P, P, P, I, I, I, I, I, I, P, P, P],
    'sed_stat_partials': [],
    'sed_colstats_f32': [P, L, I, P, P],
    'sed_bn0_aug_mix_fwd': [P, P, P, P, I, P, I, P, I, I, I, P, P],
    'sed_bn0_bwd_reduce': [P, P, P, P, P, I, P, I, P, I, I, I, P, P],
    'sed_spec_augment_f32': [P, I, I, I, I, P, I, P, I, P],
    'sed_mix_pairs_f32': [P, P, I, I, P, P],
    'sed_reduce_partials': [P, I, L, P, I, F, P],
    'sed_conv_c1_grid': [],
    'sed_conv_c1_fwd': [P, P, P, P, I, I, I, I, P],
    'sed_conv_c1_wgrad': [P, P, P, I, I, I, I, P],
    'sed_bn_apply_conv_c1_wgrad': [P, P, P, P, P, P, P, P, P, P, I, I, I, I, P],
    'sed_conv_c1_dgrad': [P, P, P, I, I, I, I, P],
    'sed_linear_partials': [],
    'sed_linear_pair_fwd': [P, P, P, I, P, P, I, L, I, P, P, P],
    'sed_linear_small_fwd': [P, P, P, L, I, I, P, P],
    'sed_linear_small_bwd': [P, P, P, L, I, I, P, I, P, P, P],
    'sed_head_pool_fwd': [P, I, I, I, I, I, P, P, P, P, P],
    'sed_head_pool_bwd': [P, P, P, I, I, I, I, P, P],
    'sed_head_att_fwd': [P, P, I, I, I, I, I, F, P, P, P, P, P],
    'sed_head_att_bwd': [P, P, P, P, P, I, I, I, I, F, P, P, P],
    'sed_bce_fwd_bwd': [P, P, L, F, P, P, P],
    'sed_gemm_tc': [P, P, P, P, L, I, I, P],
    'sed_gemm_tn_tc_splits': [L, I, I],
    'sed_gemm_tn_tc': [P, I, P, I, P, L, I, I, P],
    'sed_split_bf16x3': [P, L, I, I, P, P],
    'sed_transpose_to_bf16': [P, I, I, P, P],
    'sed_colsum_f32': [P, L, I, P, P],
    'sed_gru_workspace_bytes': [I, I, I],
    'sed_gru_fwd': [P, P, P, P, P, P, I, I, I, P],
    'sed_gru_bwd_bias_rows': [I],
    'sed_gru_bwd': [P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, P],
    'sed_attention_fwd': [P, P, P, I, I, I, I, I, I, I, F, F, U, U, P, P, P, P],
    'sed_attention_bwd': [P, P, P, I, I, I, I, I, I, I, F, F, U, U, P, P, P, P, P, P, P],
    'sed_dropout_relu_fwd': [P, L, F, U, U, P, P, P],
    'sed_dropout_relu_bwd': [P, P, L, F, P, P],
    'sed_vad_count': [P, P, I, I, I, P, P, P, P, P, P, P, P],
    'sed_vad_fill': [P, P, I, I, I, P, P, P, P, P, P, P, P],
    'sed_adam_amsgrad': [P, P, P, P, P, L, F, F, F, F, I, F, P, P],
}
_RESTYPES = {
    'sed_last_error_string': ctypes.c_char_p,
    'sed_launch_count': ctypes.c_ulonglong,
    'sed_gru_workspace_bytes': ctypes.c_longlong,
}

_lock = threading.Lock()
_lib = None


def lib():
    """The loaded library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        'libsedb200.so is missing (%s). Build it with '
                        '`python -m sound_event_detection_dcase2017_task4_b200.build`; '
                        'there is no CPU / PyTorch fallback for this path.' % LIB_PATH)
                handle = ctypes.CDLL(LIB_PATH)
                for name, argtypes in SIGNATURES.items():
                    fn = getattr(handle, name)          # AttributeError = ABI drift: fail loudly
                    fn.argtypes = argtypes
                    fn.restype = _RESTYPES.get(name, ctypes.c_int)
                _lib = handle
    return _lib


# Optional per-entry-point device timing (bench.py's roofline leg / tools): when PROFILE is a list,
# every call is bracketed by CUDA events on torch's current stream and (name, tag, start, end) is
# appended; the caller synchronises and reads elapsed times.  None (default) = no overhead.
PROFILE = None
PROFILE_TAG = None


def call(name, *args):
    fn = getattr(lib(), name)
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        status = fn(*args)
        e1.record()
        PROFILE.append((name, PROFILE_TAG, e0, e1))
    else:
        status = fn(*args)
    if status != 0:
        msg = lib().sed_last_error_string()
        raise RuntimeError('%s failed (status %d): %s' % (name, status, (msg or b'').decode()))


def launch_count():
    return int(lib().sed_launch_count())


def ptr(t):
    """Device pointer of a tensor (0 for None)."""
    return 0 if t is None else t.data_ptr()


def stream_of(t):
    import torch
    return torch.cuda.current_stream(t.device).cuda_stream
