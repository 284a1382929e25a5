"""ctypes binding of libgwn.so (the C ABI declared in include/gwn.h).

There is no CPU or eager fallback: if the shared library is missing or a call
fails, an exception is raised (``GwnError``)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GWN_LIB') or os.path.join(HERE, 'libgwn.so')     # GWN_LIB: a debug build (libgwn_trace.so)

GWN_F32, GWN_BF16 = 0, 1
MAX_SUPPORTS, MAX_TAPS, MAX_LAYERS = 4, 8, 32

vp = C.c_void_p


class GwnError(RuntimeError):
    pass


class LayerCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('N', 'V', 'Lin', 'Lout', 'Lf', 'taps', 'dilation', 'n_supports', 'order',
                                       'dtype', 'training', 'has_gconv')] + \
               [('dropout_p', C.c_float), ('seed', C.c_uint64), ('offset', C.c_uint64)]


class Ell(C.Structure):
    _fields_ = [('idx', vp * 2), ('val', vp * 2), ('width', C.c_int)]


class LayerFwdArgs(C.Structure):
    _fields_ = [('u_prev', vp), ('scale', vp), ('shift', vp), ('w_fg', vp), ('b_fg', vp), ('w_mlp', vp),
                ('b_mlp', vp), ('supports', vp * MAX_SUPPORTS), ('drop_mask', vp), ('rng', vp), ('hop_mats', vp), ('ws_w', vp), ('a', vp),
                ('b', vp), ('z_last', vp), ('u', vp), ('stats', vp), ('ws_cat', vp),
                ('bn_stats', vp), ('bn_gamma', vp), ('bn_beta', vp), ('bn_running_mean', vp), ('bn_running_var', vp),
                ('bn_mean', vp), ('bn_rstd', vp), ('bn_count', C.c_double), ('bn_momentum', C.c_float),
                ('bn_eps', C.c_float), ('ell', C.POINTER(Ell))]


class LayerBwdArgs(C.Structure):
    _fields_ = [('u_prev', vp), ('scale', vp), ('shift', vp), ('w_fg', vp), ('w_mlp', vp),
                ('supports', vp * MAX_SUPPORTS), ('support_needs_grad', C.c_int * MAX_SUPPORTS),
                ('drop_mask', vp), ('rng', vp), ('hop_mats', vp), ('ws_w', vp), ('a', vp), ('b', vp), ('du', vp), ('dz_last', vp), ('dx_prev', vp),
                ('dx_stats', vp), ('dw_fg', vp), ('db_fg', vp), ('dw_mlp', vp), ('db_mlp', vp),
                ('d_supports', vp * MAX_SUPPORTS), ('ws_cat', vp), ('ws_dcat', vp), ('ws_dfg', vp), ('outputs_zeroed', C.c_int),
                ('dx_prev_bf16', C.c_int), ('d_supports_sq', vp * MAX_SUPPORTS), ('ell', C.POINTER(Ell))]


class HeadCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('N', 'V', 'Lf', 'n_layers', 'S', 'E', 'O', 'dtype')]


class HeadFwdArgs(C.Structure):
    _fields_ = [('z_last', vp * MAX_LAYERS), ('w_skip', vp), ('b_skip', vp), ('w_end1', vp), ('b_end1', vp),
                ('w_end2', vp), ('b_end2', vp), ('s1', vp), ('e1', vp), ('out', vp), ('ws', vp)]


class HeadBwdArgs(C.Structure):
    _fields_ = [('z_last', vp * MAX_LAYERS), ('w_skip', vp), ('w_end1', vp), ('w_end2', vp), ('s1', vp),
                ('e1', vp), ('dout', vp), ('dw_skip', vp), ('db_skip', vp), ('dw_end1', vp), ('db_end1', vp),
                ('dw_end2', vp), ('db_end2', vp), ('dz_last', vp * MAX_LAYERS), ('ws_do', vp), ('ws_de1', vp),
                ('ws_ds1', vp)]


class HeadTcFwdArgs(C.Structure):
    _fields_ = [(n, vp) for n in ('zcat', 'w_skip', 'b_skip', 'w_end1', 'b_end1', 'w_end2', 'b_end2', 's1', 'e1',
                                  'out', 'ws_w')]


class HeadTcBwdArgs(C.Structure):
    _fields_ = [(n, vp) for n in ('zcat', 'w_skip', 'w_end1', 'w_end2', 's1', 'e1', 'dout', 'dw_skip', 'db_skip',
                                  'dw_end1', 'db_end1', 'dw_end2', 'db_end2')] + \
               [('dz_last', vp * MAX_LAYERS)] + [(n, vp) for n in ('ws_do', 'ws_de1', 'ws_ds1', 'ws_w')] + \
               [('outputs_zeroed', C.c_int)]


class PackCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('n_layers', 'taps', 'mlp_in', 'S', 'E', 'O', 'Opad')]


class PackPtrs(C.Structure):
    _fields_ = [(n, vp * MAX_LAYERS) for n in ('w_filter', 'b_filter', 'w_gate', 'b_gate', 'w_mlp', 'w_skip', 'b_skip')] + \
               [(n, vp) for n in ('w_end1', 'w_end2', 'b_end2')]


class UnpackPtrs(C.Structure):
    _fields_ = [(n, vp * MAX_LAYERS) for n in ('w_fg', 'b_fg', 'w_mlp')] + \
               [(n, vp) for n in ('w_skip', 'b_skip', 'w_end1', 'w_end2', 'b_end2')]


# every symbol include/gwn.h declares: name -> (restype, argtypes)
_i, _ll, _f, _d = C.c_int, C.c_longlong, C.c_float, C.c_double
SIGNATURES = {
    'gwn_last_error': (C.c_char_p, []),
    'gwn_version': (_i, []),
    'gwn_launch_count': (_ll, []),
    'gwn_check_device': (_i, []),
    'gwn_adp_fwd': (_i, [vp, vp, vp, vp, _i, _i, vp]),
    'gwn_adp_bwd': (_i, [vp, vp, vp, vp, vp, vp, vp, _i, _i, vp]),
    'gwn_adp_fwd_pair': (_i, [vp, vp, vp, _i, _i, vp]),
    'gwn_adp_pair_bwd': (_i, [vp, vp, vp, vp, vp, vp, vp, _i, _i, vp]),
    'gwn_start_fwd': (_i, [vp, vp, vp, vp, _i, _i, _i, _i, _i, _i, vp]),
    'gwn_start_bwd': (_i, [vp, vp, vp, _i, vp, vp, vp, _i, _i, _i, _i, _i, vp]),
    'gwn_start_tc_supported': (_i, [_i]),
    'gwn_start_fwd_tc': (_i, [vp, vp, vp, vp, vp, vp, _i, _i, _i, _i, _i, vp]),
    'gwn_start_bwd_tc': (_i, [vp, vp, vp, vp, vp, vp, vp, _i, _i, _i, _i, _i, vp]),
    'gwn_layer_fwd': (_i, [C.POINTER(LayerCfg), C.POINTER(LayerFwdArgs), vp]),
    'gwn_layer_bwd': (_i, [C.POINTER(LayerCfg), C.POINTER(LayerBwdArgs), vp]),
    'gwn_gcn_fwd': (_i, [vp, vp, vp, vp, vp, _i, vp, vp, vp, _f, C.c_uint64, C.c_uint64, vp, vp, _i, _i, _i, _i, vp]),
    'gwn_gcn_bwd_t_supported': (_i, [_i, _i, _i]),
    'gwn_gcn_bwd_t': (_i, [vp, vp, vp, vp, vp, _i, vp, _f, C.c_uint64, C.c_uint64, _i, vp, vp, vp, vp, vp, _i, _i, _i, _i, vp]),
    'gwn_gcn_bwd': (_i, [vp, vp, vp, vp, vp, _i, vp, vp, _f, C.c_uint64, C.c_uint64, _i, vp, vp, vp, vp, _i, _i, _i, _i, vp]),
    'gwn_dropout_apply': (_i, [vp, vp, _ll, _f, C.c_uint64, C.c_uint64, vp]),
    'gwn_bn_fold': (_i, [vp, _d, vp, vp, vp, vp, _f, _f, _i, vp, vp, vp, vp, vp]),
    'gwn_bn_bwd': (_i, [vp, _i, vp, _i, vp, _d, vp, vp, vp, _i, vp, vp, vp, _ll, vp]),
    'gwn_head_fwd': (_i, [C.POINTER(HeadCfg), C.POINTER(HeadFwdArgs), vp]),
    'gwn_head_bwd': (_i, [C.POINTER(HeadCfg), C.POINTER(HeadBwdArgs), vp]),
    'gwn_head_tc_ws_bytes': (_ll, [_i, _i, _i, _i]),
    'gwn_head_fwd_tc': (_i, [C.POINTER(HeadCfg), C.POINTER(HeadTcFwdArgs), vp]),
    'gwn_head_bwd_tc': (_i, [C.POINTER(HeadCfg), C.POINTER(HeadTcBwdArgs), vp]),
    'gwn_hop_mode': (_i, [_i, _i]),
    'gwn_hop_mats_bytes': (_i, [_i, _i]),
    'gwn_hop_mats_prep': (_i, [vp, _i, _i, vp, vp]),
    'gwn_hop_tc': (_i, [vp, _i, _i, vp, _i, _i, _i, _i, _i, vp]),
    'gwn_support_images_bytes': (_ll, [_i, _i]),
    'gwn_support_images_prep': (_i, [vp, _i, _i, vp, vp]),
    'gwn_hop_big': (_i, [vp, _i, _i, _i, vp, vp, vp, _ll, _i, vp]),
    'gwn_dadj_big': (_i, [vp, vp, vp, _ll, _i, vp]),
    'gwn_hop_ell': (_i, [vp, vp, _i, vp, vp, vp, _ll, _i, vp]),
    'gwn_mse_loss_fwd': (_i, [vp, vp, _ll, vp, vp]),
    'gwn_mse_loss_bwd': (_i, [vp, vp, vp, _ll, vp, vp]),
    'gwn_gemm_test': (_i, [vp, vp, vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, vp]),
    'gwn_pack_offsets': (_ll, [C.POINTER(PackCfg), C.POINTER(_ll)]),
    'gwn_pack_params': (_i, [C.POINTER(PackCfg), C.POINTER(PackPtrs), vp, vp]),
    'gwn_unpack_total': (_ll, [C.POINTER(PackCfg)]),
    'gwn_unpack_grads': (_i, [C.POINTER(PackCfg), C.POINTER(UnpackPtrs), vp, vp]),
    'gwn_node_mix': (_i, [vp, _i, _i, vp, _i, _i, _i, vp, _i, _i, _i, _i, vp]),
    'gwn_gather_flat': (_i, [C.POINTER(vp), C.POINTER(_ll), _i, vp, vp]),
    'gwn_peer_header_bytes': (_ll, []),
    'gwn_peer_alloc': (_i, [_ll, C.POINTER(vp)]),
    'gwn_peer_free': (_i, [vp]),
    'gwn_peer_export': (_i, [vp, vp]),
    'gwn_peer_open': (_i, [vp, C.POINTER(vp)]),
    'gwn_peer_close': (_i, [vp]),
    'gwn_adam_flat': (_i, [vp, vp, vp, vp, _ll, _f, _f, _f, _f, vp, vp]),
    'gwn_allreduce_adam': (_i, [vp, vp, vp, _ll, _f, _f, _f, _f, vp, C.POINTER(vp), _i, _i, vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libgwn.so (once).  Raises GwnError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GwnError(f'{LIB_PATH} is missing: build it with `python -m multimodal_outage_b200.build` '
                           '(there is no fallback path)')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().gwn_last_error()
        raise GwnError(f'{what} failed ({rc}): {msg.decode() if msg else "?"}')


_device_ok = set()


def require_b200(device_index: int) -> None:
    """The kernels are sm_100a-only; anything else is an error, not a fallback."""
    if device_index in _device_ok:
        return
    import torch
    with torch.cuda.device(device_index):
        check(lib().gwn_check_device(), 'gwn_check_device')
    _device_ok.add(device_index)
