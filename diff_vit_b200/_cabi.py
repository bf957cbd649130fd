"""ctypes binding of libp2vit_b200.so (C ABI declared in include/p2v.h).

The library is built in-tree by ``python -m diff_vit_b200.build``.  Loading fails loudly when it is
missing: quantized execution has no other implementation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libp2vit_b200.so')

EPI_GELU, EPI_RESIDUAL, EPI_OUT_POT, EPI_OUT_F32 = 1, 2, 4, 8

_fp = C.POINTER(C.c_float)
_i8p = C.POINTER(C.c_int8)
_vp = C.c_void_p


class Epilogue(C.Structure):
    _fields_ = [('acc_scale', _vp), ('bias', _vp), ('out_scale', _vp), ('out_rscale', _vp), ('res_scale', _vp),
                ('out2_scale', _vp), ('residual', _vp), ('aux_codes', _vp), ('out_f32', _vp),
                ('out_zp', C.c_float), ('flags', C.c_uint32)]


class LayerNorm(C.Structure):
    _fields_ = [('in_mask', _vp), ('gamma', _vp), ('beta', _vp), ('ln_out_scale', _vp), ('ln_out_rscale', _vp),
                ('post_mul', _vp), ('post_div1', _vp), ('post_div2', C.c_float), ('post_zp', C.c_float),
                ('in_scale1', C.c_float), ('pot', C.c_int), ('pre_clamp', C.c_int)]


class Attention(C.Structure):
    _fields_ = [('score_mul', C.c_float), ('score_zp', C.c_float), ('out_mul', C.c_double), ('out_zp', C.c_float),
                ('softmax_levels', C.c_int), ('exp_lut', _vp), ('dump_scores', _vp), ('dump_softmax', _vp),
                ('in_zp', C.c_float), ('lut_sig_bits', C.c_int32), ('force_legacy', C.c_int32)]


class WindowAttention(C.Structure):
    _fields_ = [('perm', _vp), ('region', _vp), ('bias', _vp), ('exp_lut', _vp), ('r3', _vp), ('exp_lut64', _vp),
                ('lut_n', C.c_int32),
                ('n', C.c_int32), ('heads', C.c_int32), ('windows', C.c_int32), ('tokens', C.c_int32),
                ('channels', C.c_int32), ('qshift', C.c_int32), ('qscale', C.c_float), ('acc_scale', C.c_double),
                ('qk_scale', C.c_float), ('err_mul', C.c_float),
                ('a1_scale', C.c_float), ('a1_rscale', C.c_float), ('a2_rscale', C.c_float), ('mask_int', C.c_int32),
                ('out_unit', C.c_float), ('out_rscale', C.c_float), ('softmax_levels', C.c_int32),
                ('dump_a1', _vp), ('dump_a2', _vp), ('dump_softmax', _vp)]


class LinearDesc(C.Structure):
    _fields_ = [('w', _vp), ('n', C.c_int32), ('k', C.c_int32), ('epi', Epilogue)]


class BlockDesc(C.Structure):
    _fields_ = [('norm1', LayerNorm), ('norm2', LayerNorm), ('qkv', LinearDesc), ('proj', LinearDesc),
                ('fc1', LinearDesc), ('fc2', LinearDesc), ('attn', Attention)]


class SwinBlockDesc(C.Structure):
    _fields_ = [('norm1', LayerNorm), ('norm2', LayerNorm), ('qkv', LinearDesc), ('proj', LinearDesc),
                ('fc1', LinearDesc), ('fc2', LinearDesc), ('attn', WindowAttention)]


class SwinStageDesc(C.Structure):
    _fields_ = [('height', C.c_int32), ('width', C.c_int32), ('dim', C.c_int32), ('depth', C.c_int32),
                ('blocks', C.POINTER(SwinBlockDesc)), ('has_merge', C.c_int32), ('merge_idx', _vp),
                ('merge_norm', LayerNorm), ('reduction', LinearDesc)]


class SwinDesc(C.Structure):
    _fields_ = [('img_size', C.c_int32), ('patch_size', C.c_int32), ('in_chans', C.c_int32), ('embed_dim', C.c_int32),
                ('num_stages', C.c_int32), ('num_classes', C.c_int32), ('input_scale', C.c_float),
                ('patch_embed', LinearDesc), ('pe_norm', LayerNorm), ('stages', C.POINTER(SwinStageDesc)),
                ('norm', LayerNorm), ('pool_in_scale', C.c_float), ('pool_out_scale', C.c_float), ('head', LinearDesc)]


class VitDesc(C.Structure):
    _fields_ = [('img_size', C.c_int32), ('patch_size', C.c_int32), ('in_chans', C.c_int32),
                ('embed_dim', C.c_int32), ('depth', C.c_int32), ('num_heads', C.c_int32),
                ('hidden_dim', C.c_int32), ('num_classes', C.c_int32),
                ('input_scale', C.c_float), ('input_zp', C.c_float),
                ('patch_embed', LinearDesc),
                ('pe_scale', C.c_float), ('pe_zp', C.c_float), ('embed_scale', C.c_float), ('embed_zp', C.c_float),
                ('cls_value', _vp), ('pos_value', _vp), ('embed_out_scale', _vp),
                ('blocks', C.POINTER(BlockDesc)),
                ('norm', LayerNorm), ('head', LinearDesc)]


# name -> (restype, argtypes); every entry point include/p2v.h declares
SYMBOLS = {
    'p2v_last_error': (C.c_char_p, []),
    'p2v_version': (C.c_int, []),
    'p2v_check_device': (C.c_int, [C.c_int]),
    'p2v_gemm_i8': (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(Epilogue), _vp]),
    'p2v_gemm_i8_simt': (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(Epilogue), _vp]),
    'p2v_gemm_set_mode': (C.c_int, [C.c_int]),
    'p2v_set_pdl': (C.c_int, [C.c_int]),
    'p2v_test_gelu_fast': (C.c_int, [C.c_float, _vp, _vp]),
    'p2v_gemm_i8_acc': (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    'p2v_quant_patchify': (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _vp]),
    'p2v_quant_patchify_u8': (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _vp,
                                        _vp, _vp]),
    'p2v_embed_assemble': (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                     _vp, _vp, _vp, _vp]),
    'p2v_layernorm_int': (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_int, C.c_int, C.POINTER(LayerNorm), _vp]),
    'p2v_attention_int': (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.POINTER(Attention), _vp]),
    'p2v_attention_tc_set_timeline': (C.c_int, [_vp]),
    'p2v_attention_tc_set_skew': (C.c_int, [C.c_int]),
    'p2v_fake_quant_f32': (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int, C.c_int64, _vp, _vp, C.c_int, C.c_int, _vp]),
    'p2v_layernorm_int_f32': (C.c_int, [_vp, _vp, C.c_int64, C.c_int, _vp, _vp, C.c_float, _vp, _vp, _vp, _vp, _vp]),
    'p2v_softmax_log_int_f32': (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_float,
                                          C.c_float, C.c_int, C.c_int, _vp]),
    'p2v_requant_eltwise': (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int, _vp, _vp, _vp, C.c_float, _vp]),
    'p2v_window_attention_int': (C.c_int, [_vp, _vp, C.c_int, C.POINTER(WindowAttention), _vp]),
    'p2v_gather_row_segments': (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    'p2v_avgpool_requant': (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _vp]),
    'p2v_swin_workspace_bytes': (C.c_int64, [C.POINTER(SwinDesc), C.c_int]),
    'p2v_swin_launches_per_forward': (C.c_int, [C.POINTER(SwinDesc)]),
    'p2v_swin_forward': (C.c_int, [C.POINTER(SwinDesc), _vp, _vp, _vp, C.c_int, _vp, _vp]),
    'p2v_swin_forward_u8': (C.c_int, [C.POINTER(SwinDesc), _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    'p2v_unpack_int4': (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    'p2v_select_histogram': (C.c_int, [_vp, C.c_int64, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp]),
    'p2v_observe_minmax': (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp]),
    'p2v_observe_scale_sse': (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int, _fp, C.c_int, C.c_float, C.c_float, _vp, _vp]),
    'p2v_vit_create': (C.c_int, [C.POINTER(VitDesc), C.c_int, C.POINTER(_vp)]),
    'p2v_vit_destroy': (None, [_vp]),
    'p2v_vit_workspace_bytes': (C.c_int64, [_vp, C.c_int]),
    'p2v_vit_forward': (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp]),
    'p2v_vit_forward_u8': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp]),
    'p2v_vit_forward_host': (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    'p2v_vit_launches_per_forward': (C.c_int, [_vp]),
    'p2v_vit_dump_layout': (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    'p2v_vit_dump_bytes': (C.c_int64, [_vp, C.c_int]),
}

_lib = None


class P2VError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise P2VError('%s is missing: run `python -m diff_vit_b200.build` (needs nvcc). Quantized '
                           'execution has no fallback implementation.' % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if os.environ.get('P2V_PDL') is not None:   # A/B hook, see p2v_set_pdl in include/p2v.h
            handle.p2v_set_pdl(int(os.environ['P2V_PDL']))
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise P2VError('libp2vit_b200 error %d: %s' % (rc, lib().p2v_last_error().decode()))


def ptr(t):
    """Device (or host) address of a tensor, None -> NULL."""
    return None if t is None else t.data_ptr()


def current_stream(device=None):
    """Raw handle of torch's current stream on `device` (default: the current device)."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream
