"""ctypes binding of ``libsga_b200.so`` (C ABI declared in ``include/sga_b200.h``).

There is no CPU fallback anywhere in the package: if the shared object is missing it is built
with nvcc (``sgaligner_b200.build``), and if that is impossible a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_LIB = None

c_f = c_float
c_i = c_int
c_p = c_void_p
c_l = c_int64


def _declare(lib):
    def sig(name, res, *args):
        fn = getattr(lib, name, None)
        if fn is None:      # caught by tests/test_abi.py, which checks every declared symbol
            return
        fn.restype = res
        fn.argtypes = list(args)

    sig('sga_last_error', c_char_p)
    sig('sga_version', c_i)
    sig('sga_device_info', c_i, POINTER(c_i), POINTER(c_i), POINTER(c_i))
    sig('sga_pointnet_fwd', c_i, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_i, c_p)
    sig('sga_pointnet_set_max_ctas', c_i, c_i)
    sig('sga_pointnet_stats_scratch_bytes', c_size_t, c_i)
    sig('sga_pointnet_fwd_stats', c_i, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_size_t, c_p)
    sig('sga_pointnet_bwd', c_i, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p,
        c_p, c_p, c_p, c_p, c_p, c_p, c_p)
    sig('sga_pointnet_bwd_mode', c_i, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p,
        c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p)
    sig('sga_pointnet_bn_moments', c_i, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p)
    sig('sga_csr_build', c_i, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p)
    sig('sga_gat_linear', c_i, c_p, c_i, c_l, c_i, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p)
    sig('sga_gat_aggregate', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_l, c_i, c_i, c_p, c_i, c_p, c_p)
    sig('sga_gat_aggregate_bwd', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_l, c_i, c_i, c_i, c_p, c_p, c_p, c_p,
        c_p, c_p, c_p)
    sig('sga_gat_linear_bwd', c_i, c_p, c_l, c_i, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p)
    sig('sga_cast_f64_f32', c_i, c_p, c_p, c_l, c_p)
    sig('sga_project_fuse_fwd', c_i, c_p, c_i, c_l, c_i, c_p, c_p, c_i, c_p, c_p, c_i, c_i, c_p, c_i, c_i, c_p)
    sig('sga_project_fuse_fwd_multi', c_i, POINTER(c_p), POINTER(c_i), POINTER(c_i), POINTER(c_p), POINTER(c_p), POINTER(c_p), c_i,
        c_l, c_i, c_p, c_i, c_p, c_p)
    sig('sga_project_fuse_bwd', c_i, c_p, c_l, c_i, c_p, c_i, c_p, c_p, c_p, c_i, c_i, c_p, c_i, c_i,
        c_p, c_p, c_p, c_p, c_p, c_size_t, c_p)
    sig('sga_match_sim', c_i, c_p, c_l, c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_p)
    sig('sga_match_topk_tc', c_i, c_p, c_l, c_i, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p)
    sig('sga_match_rank', c_i, c_p, c_l, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p)
    sig('sga_match_anchor_pos', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p)
    sig('sga_center_points', c_i, c_p, c_l, c_i, c_p, c_p, c_p)
    sig('sga_match_pair_metrics', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_p)
    sig('sga_loss_workspace_bytes', c_size_t, c_i, POINTER(c_i), c_l, c_i, c_i, c_i, c_i)
    sig('sga_loss_launch_count', c_i, c_i, POINTER(c_i), c_i, c_i, c_i)
    sig('sga_loss_set_gram_path', None, c_i)
    sig('sga_pointnet_gram_scratch_bytes', c_size_t)
    sig('sga_pointnet_bn_moments_gram', c_i, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_size_t, c_p)
    sig('sga_bn_running_update', c_i, c_p, ctypes.c_double, ctypes.c_float, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p)
    sig('sga_loss_fwd_bwd', c_i, POINTER(c_p), POINTER(c_i), c_i, c_l, c_p, c_p, c_p, c_p, c_i, c_i, c_i,
        c_p, c_p, c_f, c_p, c_i, POINTER(c_p), c_p, c_p, c_p, c_size_t, c_p)
    sig('sga_gemm_tf32x3', c_i, c_p, c_l, c_i, c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_i, c_i, c_i, c_p, c_l, c_p, c_i, c_p)
    sig('sga_adam_step', c_i, c_p, c_p, c_p, c_p, c_l, c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_p)
    sig('sga_adam_step_segments', c_i, c_p, c_p, c_p, c_p, c_l, c_p, c_i, c_p, c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_p)
    c_d = ctypes.c_double
    sig('sga_pct_point_moments', c_i, c_p, c_l, c_p, c_p)
    sig('sga_pct_affine_stats', c_i, c_p, c_p, c_i, c_p, c_p)
    sig('sga_bn_fold', c_i, c_p, c_d, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_f, c_f, c_i, c_p, c_p, c_p)
    sig('sga_pct_embed', c_i, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p)
    sig('sga_pct_pointwise', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p)
    sig('sga_pct_attn_stats', c_i, c_p, c_l, c_i, c_p, c_p)
    sig('sga_pct_attn', c_i, c_p, c_p, c_p, c_l, c_i, c_p, c_p)
    sig('sga_pct_cat_linear', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p)
    sig('sga_pct_cat_image_bytes', c_size_t, c_l, c_i)
    sig('sga_pct_cat_pack', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p)
    sig('sga_pct_pool_act', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p)
    sig('sga_bn_bwd_stats', c_i, c_p, c_p, c_p, c_p, c_p, c_f, c_f, c_l, c_i, c_p, c_p)
    sig('sga_bn_bwd_coef', c_i, c_p, c_p, c_d, c_p, c_p, c_p, c_p, c_i, c_f, c_i, c_p, c_p, c_p, c_p, c_p, c_p)
    sig('sga_bn_bwd_apply', c_i, c_p, c_p, c_p, c_p, c_p, c_f, c_f, c_p, c_p, c_p, c_l, c_i, c_p, c_p)
    sig('sga_bn_bwd_apply_absmax', c_i, c_p, c_p, c_p, c_p, c_p, c_f, c_f, c_p, c_p, c_p, c_l, c_i, c_p, c_l, c_p, c_p)
    sig('sga_pct_pow2_scale', c_i, c_p, c_p, c_l, c_l, c_f, c_p, c_p)
    sig('sga_pct_attn_bwd_dv', c_i, c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p)
    sig('sga_pct_scale_from_absmax', c_i, c_p, c_l, c_f, c_p, c_p)
    sig('sga_pct_attn_bwd_dk', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_i, c_p, c_p)
    sig('sga_pct_rowdot_scaled', c_i, c_p, c_p, c_p, c_l, c_i, c_p, c_p)
    sig('sga_pct_pointwise_scaled', c_i, c_p, c_p, c_l, c_i, c_p, c_p, c_p)
    sig('sga_pct_pointwise_scaled_absmax', c_i, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p)
    sig('sga_pct_pointwise_kv', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p)
    sig('sga_pct_scale_from_absmax_pair', c_i, c_p, c_p, c_l, c_f, c_p, c_p)
    sig('sga_pct_sa_input_grad', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_p, c_p)
    sig('sga_pct_embed_a1', c_i, c_p, c_p, c_p, c_p, c_l, c_p, c_p)
    sig('sga_pct_embed1_bwd_stats', c_i, c_p, c_p, c_p, c_p, c_p, c_l, c_p, c_p)
    sig('sga_pct_embed1_wgrad', c_i, c_p, c_p, c_d, c_p, c_p, c_p, c_p, c_p, c_p)
    sig('sga_pct_cat_dense_bwd', c_i, c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p)
    sig('sga_pct_cat_sparse_bwd_x', c_i, c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p)
    sig('sga_pct_cat_sparse_bwd_w', c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p)
    sig('sga_pct_residual', c_i, c_p, c_p, c_p, c_p, c_l, c_p, c_p)
    sig('sga_axpby_rows', c_i, c_p, c_f, c_p, c_f, c_p, c_f, c_p, c_p, c_l, c_i, c_p)
    sig('sga_pct_wgrad', c_i, c_p, c_l, POINTER(c_p), POINTER(c_i), c_i, POINTER(c_p), POINTER(c_l), c_i, c_p)
    sig('sga_wgrad_group', c_i, POINTER(c_p), POINTER(c_l), POINTER(c_i), POINTER(c_p), POINTER(c_l), POINTER(c_i), POINTER(c_p),
        POINTER(c_l), c_i, c_l, c_p)
    sig('sga_col_stats', c_i, c_p, c_l, c_i, c_p, c_p)
    sig('sga_bn_act_rows', c_i, c_p, c_p, c_p, c_p, c_f, c_l, c_i, c_p, c_p)
    sig('sga_gcn_aggregate', c_i, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p)
    sig('sga_linear_smallk', c_i, c_p, c_l, c_i, c_p, c_i, c_p, c_p)
    sig('sga_wgrad_smallk', c_i, c_p, c_p, c_l, c_i, c_i, c_p, c_p)
    sig('sga_relu_mask', c_i, c_p, c_p, c_l, c_p, c_p)
    sig('sga_colsum_rows', c_i, c_p, c_l, c_i, c_p, c_p)
    sig('sga_row_l2norm', c_i, c_p, c_l, c_i, c_f, c_p, c_p)
    sig('sga_normalize_bwd_rows', c_i, c_p, c_p, c_l, c_i, c_f, c_p, c_p)
    sig('sga_fuse_rows_fwd', c_i, POINTER(c_p), POINTER(c_i), c_i, c_p, c_l, c_p, c_i, c_p)
    sig('sga_fuse_rows_bwd', c_i, POINTER(c_p), POINTER(c_i), c_i, c_p, c_l, c_p, c_i, POINTER(c_p), c_p, c_p, c_p)
    sig('sga_nca_forward', c_i, c_p, c_i, c_f, c_f, c_f, c_p, c_p, c_p, c_p, c_p)
    sig('sga_nca_coef', c_i, c_p, c_i, c_f, c_f, c_f, c_p, c_p, c_p)
    sig('sga_selftest_umma', c_i, c_p, c_p, c_p, c_i, c_i, c_i, c_p)
    sig('sga_debug_set_trace', c_i, c_p)
    sig('sga_debug_tie_stats', c_i, c_p, c_i)


EXPORTS = ['sga_last_error', 'sga_version', 'sga_device_info', 'sga_pointnet_fwd', 'sga_pointnet_bwd', 'sga_pointnet_bwd_mode',
           'sga_pointnet_bn_moments', 'sga_pointnet_set_max_ctas', 'sga_pointnet_stats_scratch_bytes', 'sga_pointnet_fwd_stats', 'sga_csr_build', 'sga_gat_linear', 'sga_gat_aggregate',
           'sga_gat_aggregate_bwd', 'sga_gat_linear_bwd', 'sga_cast_f64_f32', 'sga_project_fuse_fwd', 'sga_project_fuse_fwd_multi', 'sga_project_fuse_bwd',
           'sga_match_sim', 'sga_match_topk_tc', 'sga_match_rank', 'sga_match_anchor_pos', 'sga_match_pair_metrics', 'sga_center_points', 'sga_loss_workspace_bytes', 'sga_loss_launch_count', 'sga_loss_set_gram_path', 'sga_bn_running_update', 'sga_pointnet_gram_scratch_bytes', 'sga_pointnet_bn_moments_gram',
           'sga_loss_fwd_bwd', 'sga_gemm_tf32x3', 'sga_adam_step', 'sga_adam_step_segments', 'sga_selftest_umma', 'sga_debug_set_trace', 'sga_debug_tie_stats',
           'sga_pct_point_moments', 'sga_pct_affine_stats', 'sga_bn_fold', 'sga_pct_embed', 'sga_pct_pointwise', 'sga_pct_attn_stats',
           'sga_pct_attn', 'sga_pct_cat_linear', 'sga_pct_cat_image_bytes', 'sga_pct_cat_pack', 'sga_pct_pool_act', 'sga_col_stats', 'sga_bn_act_rows',
           'sga_bn_bwd_stats', 'sga_bn_bwd_coef', 'sga_bn_bwd_apply', 'sga_bn_bwd_apply_absmax', 'sga_pct_attn_bwd_dv', 'sga_pct_scale_from_absmax', 'sga_pct_attn_bwd_dk', 'sga_pct_rowdot_scaled',
           'sga_pct_pointwise_scaled', 'sga_pct_pow2_scale', 'sga_pct_pointwise_scaled_absmax', 'sga_pct_pointwise_kv', 'sga_pct_scale_from_absmax_pair', 'sga_pct_sa_input_grad', 'sga_pct_embed_a1', 'sga_pct_embed1_bwd_stats', 'sga_pct_embed1_wgrad',
           'sga_pct_cat_dense_bwd', 'sga_pct_cat_sparse_bwd_x', 'sga_pct_cat_sparse_bwd_w', 'sga_pct_residual', 'sga_axpby_rows',
           'sga_wgrad_group', 'sga_pct_wgrad',
           'sga_gcn_aggregate', 'sga_linear_smallk', 'sga_wgrad_smallk', 'sga_relu_mask', 'sga_colsum_rows', 'sga_row_l2norm',
           'sga_normalize_bwd_rows', 'sga_fuse_rows_fwd', 'sga_fuse_rows_bwd', 'sga_nca_forward', 'sga_nca_coef']


def lib_path() -> str:
    from . import build
    return build.LIB


def get_lib():
    """Load (building first if needed) the CUDA library.  Never falls back to anything else."""
    global _LIB
    if _LIB is not None:
        return _LIB
    from . import build
    path = os.environ.get('SGA_LIB_PATH') or build.LIB      # SGA_LIB_PATH: A/B runs against another build of the library
    if not os.path.exists(path):
        try:
            build.build()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(
                'sgaligner_b200: libsga_b200.so is missing and could not be built with nvcc; '
                'there is no CPU or PyTorch fallback for the hot path') from e
    lib = ctypes.CDLL(path)
    _declare(lib)
    _LIB = lib
    return lib


def check(rc: int, what: str = ''):
    if rc != 0:
        msg = get_lib().sga_last_error().decode(errors='replace')
        raise RuntimeError(f'libsga_b200 {what} failed (code {rc}): {msg}')
