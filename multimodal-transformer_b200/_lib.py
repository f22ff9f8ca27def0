"""ctypes binding of libmt_b200.so (the C ABI declared in include/mt_b200.h).

There is NO fallback: if the library is missing, or a tensor handed to it is not a contiguous CUDA tensor of the
expected dtype, a RuntimeError is raised.  PyTorch is used for device memory and streams only.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_size_t, c_uint32, c_uint64, c_void_p)

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmt_b200.so')

MT_F32, MT_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
MT_MAX_MODS = 4


class MtEncoderCfg(Structure):
    _fields_ = [('B', c_int), ('T', c_int), ('d', c_int), ('h', c_int), ('dff', c_int), ('n_layers', c_int),
                ('dtype', c_int), ('training', c_int), ('p_drop', c_float), ('seed', c_uint64), ('stack_id', c_int),
                ('y_f32', c_int), ('grid_share', c_int), ('key_len', c_void_p)]


class MtMfnCfg(Structure):
    _fields_ = [('B', c_int), ('T', c_int), ('n_mods', c_int), ('in_dim', c_int * MT_MAX_MODS), ('hid', c_int * MT_MAX_MODS),
                ('mem_dim', c_int), ('h_att1', c_int), ('h_att2', c_int), ('h_gamma', c_int), ('h_out', c_int),
                ('dtype', c_int), ('training', c_int), ('p_gamma', c_float), ('p_out', c_float), ('seed', c_uint64), ('bwd_phase', c_int)]


class MtLstmHeadCfg(Structure):
    _fields_ = [('B', c_int), ('T', c_int), ('E', c_int), ('Hd', c_int), ('dtype', c_int), ('training', c_int)]


class MtWindowCnnCfg(Structure):
    _fields_ = [('dtype', c_int), ('n_win', c_int), ('K', c_int), ('D', c_int), ('E', c_int), ('k', c_int), ('stages', c_int),
                ('training', c_int), ('dropout_p', c_float), ('seed', c_uint64), ('site', c_uint32)]


P = c_void_p
_PROTOS = {
    # name: (restype, argtypes)
    'mt_error_string': (c_char_p, [c_int]),
    'mt_last_cuda_error': (c_char_p, []),
    'mt_check_device': (c_int, [c_int]),
    'mt_version': (c_int, []),
    'mt_launch_count': (c_uint64, []),
    'mt_prof_start': (c_int, [c_int, P]),
    'mt_prof_stop': (c_int, []),
    'mt_prof_get': (c_int, [c_int, c_char_p, c_int, POINTER(c_float), POINTER(ctypes.c_double), POINTER(ctypes.c_double)]),
    'mt_linear_fwd': (c_int, [c_int, c_int, c_int, c_int, P, c_int, P, P, P, c_int, c_int, P, c_float, c_uint64, c_uint32, P, c_size_t, P]),
    'mt_linear_ws_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_float]),
    'mt_linear_bwd': (c_int, [c_int, c_int, c_int, c_int, P, c_int, P, P, c_int, P, c_int, c_int, P, c_float, c_uint64, c_uint32, P, P, P,
                              P, c_size_t, P]),
    'mt_linear_bwd_ws_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_float]),
    'mt_layernorm_fwd': (c_int, [c_int, c_int, c_int, P, P, P, c_float, P, c_int, P]),
    'mt_layernorm_bwd': (c_int, [c_int, c_int, c_int, P, P, c_float, P, c_int, P, P, P, P, P]),
    'mt_attention_fwd': (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P, c_float, c_uint64, c_uint32, P]),
    'mt_attention_ragged_fwd': (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P, P]),
    'mt_attention_bwd': (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, c_float, c_uint64, c_uint32, P, c_size_t, P]),
    'mt_attention_bwd_ws_bytes': (c_size_t, [c_int, c_int, c_int]),
    'mt_attention_force_ffma': (c_int, [c_int]),
    'mt_attention_force_tiled': (c_int, [c_int]),
    'mt_attention_tc_fwd': (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, c_float, c_uint64, c_uint32, P, P]),
    'mt_attention_tc_bwd': (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P, P, c_float, c_uint64, c_uint32, P, P, c_size_t, P]),
    'mt_attention_tc_bwd_ws_bytes': (c_size_t, [c_int, c_int, c_int]),
    'mt_attention_force_no_tc': (c_int, [c_int]),
    'mt_attention_probs': (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    'mt_encoder_param_count': (c_size_t, [c_int, c_int, c_int]),
    'mt_encoder_ws_bytes': (c_size_t, [POINTER(MtEncoderCfg)]),
    'mt_encoder_fwd': (c_int, [POINTER(MtEncoderCfg), P, P, P, P, P, P, c_size_t, P]),
    'mt_encoder_bwd': (c_int, [POINTER(MtEncoderCfg), P, P, P, P, P, P, P, P, c_size_t, P]),
    'mt_encoder_group_ws_bytes': (c_size_t, [POINTER(MtEncoderCfg), c_int]),
    'mt_encoder_group_fwd': (c_int, [POINTER(MtEncoderCfg), c_int, POINTER(c_uint64), POINTER(c_int), P, P, c_size_t, P, P, P, P, c_size_t, P]),
    'mt_encoder_group_bwd': (c_int, [POINTER(MtEncoderCfg), c_int, POINTER(c_uint64), POINTER(c_int), P, P, c_size_t, P, P, P, P, P, P, c_size_t, P]),
    'mt_encoder_stack_fwd': (c_int, [POINTER(MtEncoderCfg), P, P, P, P, P, P, c_size_t, P]),
    'mt_encoder_stack_bwd': (c_int, [POINTER(MtEncoderCfg), P, P, P, P, P, P, P, P, c_size_t, P]),
    'mt_comm_available': (c_int, []),
    'mt_comm_unique_id': (c_int, [c_char_p]),
    'mt_comm_init': (c_int, [c_char_p, c_int, c_int, POINTER(c_void_p)]),
    'mt_comm_destroy': (c_int, [P]),
    'mt_allreduce_grads': (c_int, [P, POINTER(c_void_p), POINTER(c_size_t), c_int, P]),
    'mt_comm_overlap_arm': (c_int, [P, c_int]),
    'mt_comm_overlap_join': (c_int, [P, POINTER(c_void_p), POINTER(c_size_t), POINTER(c_int)]),
    'mt_mfn_param_count': (c_size_t, [POINTER(MtMfnCfg)]),
    'mt_mfn_ws_bytes': (c_size_t, [POINTER(MtMfnCfg)]),
    'mt_mfn_fwd': (c_int, [POINTER(MtMfnCfg), P, P, POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64), P, P, P, P, P, P, c_size_t, P]),
    'mt_mfn_bwd': (c_int, [POINTER(MtMfnCfg), P, P, POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64), P, P, POINTER(c_void_p), P, P,
                           c_size_t, P]),
    'mt_lstm_head_param_count': (c_size_t, [POINTER(MtLstmHeadCfg)]),
    'mt_lstm_head_ws_bytes': (c_size_t, [POINTER(MtLstmHeadCfg)]),
    'mt_lstm_head_fwd': (c_int, [POINTER(MtLstmHeadCfg), P, P, P, P, P, P, c_size_t, P]),
    'mt_lstm_head_bwd': (c_int, [POINTER(MtLstmHeadCfg), P, P, P, P, P, P, P, P, c_size_t, P]),
    'mt_lstm_head_force_ffma': (c_int, [c_int]),
    'mt_window_cnn_ws_bytes': (c_size_t, [POINTER(MtWindowCnnCfg)]),
    'mt_window_cnn_fwd': (c_int, [POINTER(MtWindowCnnCfg), P, P, P, P, P, P, P, P, P, c_size_t, P]),
    'mt_window_cnn_bwd': (c_int, [POINTER(MtWindowCnnCfg), P, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    'mt_ccc_batched': (c_int, [P, P, P, c_int, c_int, P, P, P, P]),
    'mt_batch_gather': (c_int, [P, c_size_t, P, c_int, c_size_t, P, P]),
    'mt_length_mask': (c_int, [P, c_int, c_int, P, P]),
    'mt_residual_dropout_fwd': (c_int, [P, P, P, c_size_t, c_float, c_uint64, c_uint32, P]),
    'mt_dropout_bwd': (c_int, [P, P, c_size_t, c_float, c_uint64, c_uint32, P]),
    'mt_cast_f32_to_bf16': (c_int, [P, P, c_size_t, P]),
    'mt_concat_fwd': (c_int, [c_int, c_int, POINTER(c_void_p), POINTER(c_int), POINTER(c_int), P, c_int, c_int, P]),
    'mt_concat_bwd': (c_int, [c_int, c_int, POINTER(c_void_p), POINTER(c_int), POINTER(c_int), P, c_int, c_int, P]),
    'mt_cast_bf16_to_f32': (c_int, [P, P, c_size_t, P]),
    'mt_mse_loss_fwd_bwd': (c_int, [P, P, c_size_t, c_float, P, P, P]),
    'mt_adam_step': (c_int, [P, P, P, P, c_size_t, c_float, c_float, c_float, c_float, c_float, c_int, P]),
    'mt_mse_loss_fwd_bwd_dev': (c_int, [P, P, c_size_t, P, P, P, P]),
    'mt_adam_step_dev': (c_int, [P, P, P, P, c_size_t, P, c_float, c_float, c_float, c_float, c_float, P, P, P]),
    'mt_set_seed_offset_ptr': (c_int, [P]),
    'mt_spin': (c_int, [c_float, P]),
    'mt_gemm': (c_int, [c_int, c_int, c_int, c_int, P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, P, c_int, c_int, P]),
    'mt_gemm_rs': (c_int, [c_int, c_int, c_int, c_int, P, P, c_int, P, c_int, P, c_int, c_float, c_uint64, c_uint32, P, c_float, P, P, P, P, P, P]),
    'mt_gemm_rs_trace': (c_int, [P]),
    'mt_gemm_engine': (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    'mt_gemm_force_simt': (c_int, [c_int]),
    'mt_mfn_force_ffma': (c_int, [c_int]),
    'mt_gemm_debug_trace': (c_int, [P]),
    'mt_gemm_tc_mode': (c_int, [c_int]),
    'mt_tune': (c_int, [c_int, c_int]),
}

EXPORTS = tuple(_PROTOS)      # every symbol include/mt_b200.h declares

_lib = None


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                               '(there is no CPU / PyTorch fallback for this path)')
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            f = getattr(L, name)          # AttributeError here means the header and the library disagree
            f.restype = res
            f.argtypes = args
        # MT_B200_TUNE="key=value,..." presets mt_tune knobs for the whole process (A/B runs of the test suite, e.g. "3=0")
        for kv in filter(None, os.environ.get('MT_B200_TUNE', '').split(',')):
            k_, v_ = kv.split('=')
            L.mt_tune(int(k_), int(v_))
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        L = lib()
        msg = L.mt_error_string(rc).decode()
        if rc == 3:
            msg += ': ' + L.mt_last_cuda_error().decode()
        raise RuntimeError(f'libmt_b200: {msg}')


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require(t, dtype=None, name='tensor'):
    """Validate a tensor handed to the library."""
    if not t.is_cuda:
        raise RuntimeError(f'{name}: the B200 path needs CUDA tensors (got {t.device}); there is no CPU fallback -- '
                           'the reference classes are the CPU implementation')
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f'{name}: expected {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise RuntimeError(f'{name}: must be contiguous')
    return t


_device_checked = set()


def check_device(dev_index):
    if dev_index not in _device_checked:
        check(lib().mt_check_device(dev_index))
        _device_checked.add(dev_index)
