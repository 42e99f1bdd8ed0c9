"""ctypes binding of the C-ABI shared library (include/ld_b200.h).

The CUDA library IS the product: importing this module without a built `libld_b200.so` raises,
there is no CPU or PyTorch fallback anywhere in `ld_b200`.  PyTorch is only used for device
memory, streams and `torch.distributed`; tensors cross the boundary as raw device pointers.
"""
import ctypes as C
import os

MAX_LEVELS = 8
NUM_TERMS = 9
F32, BF16 = 0, 1
IM_FINEGRAINED, IM_CENTER = 0, 1

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libld_b200.so')


class Pyramid(C.Structure):
    _fields_ = [('num_levels', C.c_int), ('num_imgs', C.c_int), ('h', C.c_int * MAX_LEVELS),
                ('w', C.c_int * MAX_LEVELS), ('stride', C.c_int * MAX_LEVELS)]

    @staticmethod
    def make(num_imgs, sizes, strides):
        p = Pyramid()
        assert len(sizes) == len(strides) <= MAX_LEVELS
        p.num_levels, p.num_imgs = len(sizes), int(num_imgs)
        for i, ((h, w), s) in enumerate(zip(sizes, strides)):
            p.h[i], p.w[i], p.stride[i] = int(h), int(w), int(s)
        return p

    @property
    def level_anchors(self):
        return [self.h[i] * self.w[i] for i in range(self.num_levels)]

    @property
    def rows(self):
        return self.num_imgs * sum(self.level_anchors)


class LossCfg(C.Structure):
    _fields_ = [('num_classes', C.c_int), ('reg_max', C.c_int), ('feat_channels', C.c_int),
                ('qfl_weight', C.c_float), ('qfl_beta', C.c_float),
                ('giou_weight', C.c_float), ('giou_eps', C.c_float),
                ('dfl_weight', C.c_float),
                ('ld_weight', C.c_float), ('ld_T', C.c_float),
                ('ld_vlr_weight', C.c_float), ('ld_vlr_T', C.c_float),
                ('kd_weight', C.c_float), ('kd_T', C.c_float),
                ('im_weight', C.c_float)]


class ConvDesc(C.Structure):
    _fields_ = [('in_', C.c_void_p), ('N', C.c_int), ('H', C.c_int), ('W', C.c_int), ('Cin', C.c_int),
                ('in_ld', C.c_int), ('weight', C.c_void_p), ('Cout', C.c_int), ('w_k', C.c_int),
                ('ntaps', C.c_int), ('tap_dx', C.c_int * 9), ('tap_dy', C.c_int * 9), ('tap_wk', C.c_int * 9),
                ('sx', C.c_int), ('sy', C.c_int), ('HO', C.c_int), ('WO', C.c_int), ('flat', C.c_int),
                ('out', C.c_void_p), ('OH', C.c_int), ('OW', C.c_int), ('oys', C.c_int), ('oyo', C.c_int),
                ('oxs', C.c_int), ('oxo', C.c_int), ('out_ld', C.c_int), ('out_f32', C.c_int),
                ('scale', C.c_void_p), ('bias', C.c_void_p), ('add_src', C.c_void_p), ('add_H', C.c_int),
                ('add_W', C.c_int), ('add_ld', C.c_int), ('mask_src', C.c_void_p), ('mask_ld', C.c_int),
                ('relu', C.c_int), ('accumulate', C.c_int), ('split', C.c_int), ('w_plane', C.c_int),
                ('add_f32', C.c_int), ('mask_f32', C.c_int), ('gn_stats', C.c_void_p),
                ('split_pass', C.c_int), ('pre_src', C.c_void_p)]


class GnSeg(C.Structure):
    _fields_ = [('x', C.c_void_p), ('y', C.c_void_p), ('dy', C.c_void_p), ('out', C.c_void_p), ('stats', C.c_void_p),
                ('red_ws', C.c_void_p), ('red_zeroed', C.c_int), ('N', C.c_int), ('HW', C.c_int)]


class PrepImage(C.Structure):
    _fields_ = [('src_off', C.c_longlong), ('sh', C.c_int), ('sw', C.c_int), ('dh', C.c_int), ('dw', C.c_int),
                ('flip', C.c_int)]


class WgradDesc(C.Structure):
    _fields_ = [('dy', C.c_void_p), ('N', C.c_int), ('HO', C.c_int), ('WO', C.c_int), ('Cout', C.c_int),
                ('dy_ld', C.c_int), ('x', C.c_void_p), ('H', C.c_int), ('W', C.c_int), ('Cin', C.c_int),
                ('x_ld', C.c_int), ('ntaps', C.c_int), ('tap_dx', C.c_int * 9), ('tap_dy', C.c_int * 9),
                ('tap_wk', C.c_int * 9), ('sx', C.c_int), ('sy', C.c_int), ('dw', C.c_void_p), ('w_k', C.c_int),
                ('scale', C.c_void_p), ('w_master', C.c_void_p), ('d_scale', C.c_void_p), ('split', C.c_int), ('split_pass', C.c_int)]


_SIGS = {
    'ldb200_last_error': (C.c_char_p, []),
    'ldb200_version': (C.c_int, []),
    'ldb200_check_device': (C.c_int, []),
    'ldb200_targets_workspace_bytes': (C.c_size_t, [C.POINTER(Pyramid), C.c_int]),
    'ldb200_ld_targets': (C.c_int, [C.POINTER(Pyramid)] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 11),
    'ldb200_ld_loss_prepass': (C.c_int, [C.POINTER(Pyramid), C.POINTER(LossCfg)] + [C.c_void_p] * 9),
    'ldb200_ld_loss_fused': (C.c_int, [C.POINTER(Pyramid), C.POINTER(LossCfg)] + [C.c_void_p] * 6 +
                             [C.c_int] + [C.c_void_p] * 9 + [C.c_void_p] + [C.c_void_p] * 4 +
                             [C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ldb200_ld_loss_prepass_v2': (C.c_int, [C.POINTER(Pyramid), C.POINTER(LossCfg)] + [C.c_void_p] * 6 + [C.c_int]
                                  + [C.c_void_p] * 3),
    'ldb200_ld_loss_fused_v2': (C.c_int, [C.POINTER(Pyramid), C.POINTER(LossCfg)] + [C.c_void_p] * 7 + [C.c_int]
                                + [C.c_void_p] * 9 + [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ldb200_gi_region_v2': (C.c_int, [C.POINTER(Pyramid), C.POINTER(LossCfg)] + [C.c_void_p] * 4 + [C.c_float, C.c_int]
                            + [C.c_void_p] * 5),
    'ldb200_gi_region_workspace_bytes': (C.c_size_t, [C.POINTER(Pyramid), C.c_int]),
    'ldb200_gi_region': (C.c_int, [C.POINTER(Pyramid), C.POINTER(LossCfg)] + [C.c_void_p] * 4 + [C.c_float, C.c_int]
                         + [C.c_void_p] * 5),
    'ldb200_dgqp_forward': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 6 + [C.c_void_p] * 7),
    'ldb200_dgqp_backward': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_int] * 6 + [C.c_void_p] * 7
                             + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5),
    'ldb200_scatter_add_rows': (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ldb200_conv2d_tc': (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    'ldb200_conv2d_tc_multi': (C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_void_p]),
    'ldb200_conv2d_wgrad_tc': (C.c_int, [C.POINTER(WgradDesc), C.c_void_p]),
    'ldb200_stem_im2col': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ldb200_split_bf16': (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    'ldb200_deform_im2col': (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 14 + [C.c_void_p]),
    'ldb200_stem_tc': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 7),
    'ldb200_maxpool3x3s2': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    'ldb200_gn_stats': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    'ldb200_gn_apply': (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_float, C.c_int, C.c_int, C.c_void_p]),
    'ldb200_gn_relu_bwd': (C.c_int, [C.c_void_p] * 9 + [C.c_int] * 4 + [C.c_float, C.c_int, C.c_void_p]),
    'ldb200_gn_apply_multi': (C.c_int, [C.POINTER(GnSeg), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                        C.c_int, C.c_int, C.c_void_p]),
    'ldb200_gn_relu_bwd_multi': (C.c_int, [C.POINTER(GnSeg), C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_float,
                                                                                          C.c_int, C.c_void_p]),
    'ldb200_upsample_add_bwd': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]),
    'ldb200_relu_bwd': (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int, C.c_void_p]),
    'ldb200_colsum': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ldb200_sgd_flat': (C.c_int, [C.c_void_p] * 3 + [C.c_int64] + [C.c_float] * 4 + [C.c_int, C.c_void_p]),
    'ldb200_sgd_flat_dev': (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_void_p, C.c_void_p]),
    'ldb200_bn_fold': (C.c_int, [C.c_void_p] * 4 + [C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'ldb200_bn_fold_bwd': (C.c_int, [C.c_void_p] * 4 + [C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'ldb200_mse_fwd_bwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ldb200_preprocess_batch': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                          C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_void_p]),
    'ldb200_get_bboxes_workspace_bytes': (C.c_size_t, [C.POINTER(Pyramid), C.c_int, C.c_int]),
    'ldb200_gfl_get_bboxes': (C.c_int, [C.POINTER(Pyramid), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int] + [C.c_void_p] * 5),
    'ldb200_qfl_fwd_bwd': (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.c_float, C.c_int, C.c_float] + [C.c_void_p] * 4),
    'ldb200_dfl_fwd_bwd': (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int, C.c_float] + [C.c_void_p] * 4),
    'ldb200_giou_fwd_bwd': (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_float, C.c_float] + [C.c_void_p] * 4),
    'ldb200_kl_dense_fwd_bwd': (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int, C.c_float, C.c_float,
                                          C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


class LdB200Error(RuntimeError):
    pass


def exported_symbols():
    """Every entry point `include/ld_b200.h` declares (kept in sync by tests/test_capi_symbols.py)."""
    return sorted(_SIGS)


def load():
    """Load the shared library; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise LdB200Error(
            f'{_LIB_PATH} is missing: ld_b200 has no CPU / PyTorch fallback. Build the CUDA library '
            f'first (`python -m ld_b200.build` or `__graft_entry__.build()`).')
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


# kernels launched by one call of each entry point (memsets not counted); `launch_count` is what
# bench.py reports as "gpu_launches"
_LAUNCHES = {'ldb200_ld_targets': 3, 'ldb200_ld_loss_prepass': 1, 'ldb200_ld_loss_fused': 1,
             'ldb200_scatter_add_rows': 1, 'ldb200_gi_region': 3, 'ldb200_gi_region_v2': 3, 'ldb200_ld_loss_prepass_v2': 1,
             'ldb200_ld_loss_fused_v2': 1, 'ldb200_dgqp_forward': 1, 'ldb200_dgqp_backward': 1, 'ldb200_conv2d_tc': 1, 'ldb200_conv2d_tc_multi': 1, 'ldb200_conv2d_wgrad_tc': 1,
             'ldb200_stem_im2col': 1, 'ldb200_split_bf16': 1, 'ldb200_deform_im2col': 1, 'ldb200_stem_tc': 1, 'ldb200_maxpool3x3s2': 1, 'ldb200_gn_stats': 1, 'ldb200_gn_apply': 1,
             'ldb200_gn_relu_bwd': 2, 'ldb200_gn_apply_multi': 1, 'ldb200_gn_relu_bwd_multi': 2, 'ldb200_upsample_add_bwd': 1, 'ldb200_relu_bwd': 1, 'ldb200_colsum': 1,
             'ldb200_sgd_flat': 1, 'ldb200_sgd_flat_dev': 1, 'ldb200_bn_fold': 1, 'ldb200_bn_fold_bwd': 1, 'ldb200_mse_fwd_bwd': 1, 'ldb200_kl_dense_fwd_bwd': 1,
             'ldb200_gfl_get_bboxes': 4, 'ldb200_preprocess_batch': 1, 'ldb200_qfl_fwd_bwd': 1, 'ldb200_dfl_fwd_bwd': 1, 'ldb200_giou_fwd_bwd': 1}
launch_count = {'n': 0}


def check(code, what=''):
    launch_count['n'] += _LAUNCHES.get(what, 0)
    if code != 0:
        msg = load().ldb200_last_error().decode(errors='replace')
        raise LdB200Error(f'{what} failed with code {code}: {msg}')


def ptr(t):
    """Device (or host) pointer of a contiguous torch tensor, None -> NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), 'ld_b200 kernels need contiguous tensors'
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
