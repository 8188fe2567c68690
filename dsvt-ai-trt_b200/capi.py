"""ctypes binding of include/dsvt_b200.h for torch CUDA tensors.

PyTorch is plumbing here (device memory, streams); every function launches the hand-written
sm_100a kernels through the C ABI and raises if the call fails.  No fallback path exists.
"""
import ctypes
from ctypes import POINTER, Structure, c_float, c_int32, c_size_t, c_uint64, c_void_p

import torch

from ._lib import load_library

DSVT_ATTN_FP32, DSVT_ATTN_FP16, DSVT_ATTN_FP32_TC, DSVT_ATTN_FP16_GEMM = 0, 2, 3, 4
DSVT_LINEAR_TF32 = 1          # dense linear layers only (tc_linear.cu)


class P2FParams(Structure):
    _fields_ = [("batch", c_int32), ("max_points_num", c_int32), ("max_points_num_voxel_filter", c_int32),
                ("max_pillars_num", c_int32), ("point_feature_num", c_int32), ("feature_num", c_int32),
                ("max_num_points_per_voxel", c_int32),
                ("x_min", c_float), ("x_max", c_float), ("y_min", c_float), ("y_max", c_float),
                ("z_min", c_float), ("z_max", c_float),
                ("voxel_x", c_float), ("voxel_y", c_float), ("voxel_z", c_float),
                ("grid_x", c_int32), ("grid_y", c_int32), ("grid_z", c_int32), ("zero_tails", c_int32)]


class WPParams(Structure):
    _fields_ = [("batch", c_int32), ("max_pillars_num", c_int32), ("max_win_num", c_int32),
                ("max_voxel_num_per_win", c_int32),
                ("sparse_shape_x", c_int32), ("sparse_shape_y", c_int32), ("sparse_shape_z", c_int32),
                ("win_shape_x", c_int32), ("win_shape_y", c_int32), ("win_shape_z", c_int32),
                ("shift_x", c_int32), ("shift_y", c_int32), ("shift_z", c_int32), ("zero_tails", c_int32)]


class GSParams(Structure):
    _fields_ = [("batch", c_int32), ("voxel_num_set", c_int32), ("max_win_num", c_int32),
                ("max_voxel_num_per_win", c_int32),
                ("win_shape_x", c_int32), ("win_shape_y", c_int32), ("win_shape_z", c_int32),
                ("num_heads", c_int32), ("zero_tails", c_int32)]


class GeluParams(Structure):
    _fields_ = [("batch", c_int32), ("max_pillars_num", c_int32), ("channel_num", c_int32), ("zero_tails", c_int32)]


class LNParams(Structure):
    _fields_ = [("batch", c_int32), ("max_pillars_num", c_int32), ("channel_num", c_int32), ("eps", c_float),
                ("zero_tails", c_int32)]


class FBParams(Structure):
    _fields_ = [("batch", c_int32), ("max_top_k", c_int32),
                ("x_min", c_float), ("x_max", c_float), ("y_min", c_float), ("y_max", c_float),
                ("z_min", c_float), ("z_max", c_float),
                ("voxel_x", c_float), ("voxel_y", c_float), ("voxel_z", c_float),
                ("score_threshold", c_float), ("zero_tails", c_int32)]


class AttnParams(Structure):
    _fields_ = [("batch", c_int32), ("max_set_num", c_int32), ("voxel_num_set", c_int32), ("channel_num", c_int32),
                ("num_heads", c_int32), ("max_pillars_num", c_int32), ("axis_id", c_int32), ("precision", c_int32),
                ("zero_tails", c_int32)]


_sig_done = False


def _lib():
    global _sig_done
    lib = load_library()
    if not _sig_done:
        lib.dsvt_last_error.restype = ctypes.c_char_p
        lib.dsvt_launch_count.restype = c_uint64
        lib.dsvt_points2features_workspace_size.restype = c_size_t
        lib.dsvt_window_partition_workspace_size.restype = c_size_t
        lib.dsvt_get_set_workspace_size.restype = c_size_t
        lib.dsvt_set_attention_workspace_size.restype = c_size_t
        lib.dsvt_attention_weights_create.restype = c_void_p
        lib.dsvt_attention_weights_create.argtypes = [c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
        lib.dsvt_attention_weights_destroy.argtypes = [c_void_p]
        lib.dsvt_linear_weights_create.restype = c_void_p
        lib.dsvt_linear_weights_create.argtypes = [c_int32, c_int32, c_void_p, c_void_p, c_int32]
        lib.dsvt_linear_weights_destroy.argtypes = [c_void_p]
        lib.dsvt_small_linear_create.restype = c_void_p
        lib.dsvt_vfe_fused_workspace_size.restype = c_size_t
        lib.dsvt_small_linear_create.argtypes = [c_int32, c_int32, c_void_p, c_void_p, c_void_p]
        lib.dsvt_small_linear_destroy.argtypes = [c_void_p]
        _sig_done = True
    return lib


class DsvtError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise DsvtError(f"{what} failed with code {rc}: {_lib().dsvt_last_error().decode()}")


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(_lib().dsvt_launch_count())


def _need(t, dtype, name):
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise DsvtError(f"{name}: expected a contiguous CUDA tensor of {dtype}")


# ----------------------------------------------------------------------------------------------
def p2f_params(cfg, batch=1, zero_tails=1):
    return P2FParams(batch, cfg.max_points_num, cfg.max_points_num_voxel_filter, cfg.max_pillars_num,
                     cfg.point_feature_num, cfg.feature_num, cfg.max_num_points_per_voxel,
                     cfg.x_min, cfg.x_max, cfg.y_min, cfg.y_max, cfg.z_min, cfg.z_max,
                     cfg.voxel_x, cfg.voxel_y, cfg.voxel_z, cfg.grid_x, cfg.grid_y, cfg.grid_z, zero_tails)


class Points2Features:
    """Voxeliser with pre-allocated outputs / workspace (reference: Points2FeaturesPlugin)."""

    def __init__(self, cfg, batch=1, device="cuda", zero_tails=1):
        self.cfg, self.batch = cfg, batch
        self.p = p2f_params(cfg, batch, zero_tails)
        lib = _lib()
        self.ws_bytes = int(lib.dsvt_points2features_workspace_size(ctypes.byref(self.p)))
        if self.ws_bytes == 0:
            raise DsvtError("points2features: " + lib.dsvt_last_error().decode())
        B = batch
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=device)
        self.point_features = torch.empty(B, cfg.max_points_num_voxel_filter, 10, dtype=torch.float32, device=device)
        self.point_index_in_voxel = torch.empty(B, cfg.max_pillars_num, cfg.max_num_points_per_voxel,
                                                dtype=torch.int32, device=device)
        self.coords = torch.empty(B, cfg.max_pillars_num, 4, dtype=torch.int32, device=device)
        self.point_num_in_voxel = torch.empty(B, cfg.max_pillars_num, dtype=torch.int32, device=device)
        self.pillar_num = torch.empty(B, dtype=torch.int32, device=device)
        self.point_num = torch.empty(B, dtype=torch.int32, device=device)

    def __call__(self, points, points_size):
        _need(points, torch.float32, "points")
        _need(points_size, torch.int32, "points_size")
        rc = _lib().dsvt_points2features_launch(
            ctypes.byref(self.p), _ptr(points), _ptr(points_size), _ptr(self.point_features),
            _ptr(self.point_index_in_voxel), _ptr(self.coords), _ptr(self.point_num_in_voxel),
            _ptr(self.pillar_num), _ptr(self.point_num), _ptr(self.ws), c_size_t(self.ws_bytes), _stream())
        _check(rc, "dsvt_points2features_launch")
        return self


class WindowPartition:
    def __init__(self, cfg, which, batch=1, device="cuda", zero_tails=1):
        wx, wy, wz = cfg.win_shapes[which]
        sx, sy, sz = cfg.shifts[which]
        self.cfg, self.batch = cfg, batch
        self.p = WPParams(batch, cfg.max_pillars_num, cfg.max_win_num, cfg.max_voxel_num_per_win,
                          cfg.grid_x, cfg.grid_y, cfg.grid_z, wx, wy, wz, sx, sy, sz, zero_tails)
        lib = _lib()
        self.ws_bytes = int(lib.dsvt_window_partition_workspace_size(ctypes.byref(self.p)))
        if self.ws_bytes == 0:
            raise DsvtError("window_partition: " + lib.dsvt_last_error().decode())
        B, mw, mv, mp = batch, cfg.max_win_num, cfg.max_voxel_num_per_win, cfg.max_pillars_num
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=device)
        self.global_index = torch.empty(B, mw, mv, dtype=torch.int32, device=device)
        self.coors_in_win = torch.empty(B, mw, mv, 3, dtype=torch.int32, device=device)
        self.voxel_num_in_win = torch.empty(B, mw, dtype=torch.int32, device=device)
        self.win_num = torch.empty(B, dtype=torch.int32, device=device)
        self.coors_in_win_2d = torch.empty(B, mp, 3, dtype=torch.int32, device=device)
        self.coors_in_win_x_y = torch.empty(B, mp, 2, dtype=torch.float32, device=device)

    def __call__(self, coords, voxel_num):
        _need(coords, torch.int32, "coords")
        _need(voxel_num, torch.int32, "voxel_num")
        rc = _lib().dsvt_window_partition_launch(
            ctypes.byref(self.p), _ptr(coords), _ptr(voxel_num), _ptr(self.global_index), _ptr(self.coors_in_win),
            _ptr(self.voxel_num_in_win), _ptr(self.win_num), _ptr(self.coors_in_win_2d), _ptr(self.coors_in_win_x_y),
            _ptr(self.ws), c_size_t(self.ws_bytes), _stream())
        _check(rc, "dsvt_window_partition_launch")
        return self


class GetSet:
    def __init__(self, cfg, which, batch=1, device="cuda", zero_tails=1):
        wx, wy, wz = cfg.win_shapes[which]
        self.cfg, self.batch = cfg, batch
        self.p = GSParams(batch, cfg.voxel_num_set, cfg.max_win_num, cfg.max_voxel_num_per_win, wx, wy, wz,
                          cfg.num_heads, zero_tails)
        B, mw, S, H = batch, cfg.max_win_num, cfg.voxel_num_set, cfg.num_heads
        self.global_index_in_set = torch.empty(B, 2, mw, S, dtype=torch.int32, device=device)
        self.set_voxel_mask = torch.empty(B, 2, mw, S, dtype=torch.float32, device=device)
        self.set_num = torch.empty(B, dtype=torch.int32, device=device)
        self.mask_expand_0 = torch.empty(B, mw, H, S, dtype=torch.float32, device=device)
        self.mask_expand_1 = torch.empty(B, mw, H, S, dtype=torch.float32, device=device)

    def __call__(self, global_index, coors_in_win, voxel_num_in_win, win_num):
        for t, n in ((global_index, "global_index"), (coors_in_win, "coors_in_win"),
                     (voxel_num_in_win, "voxel_num_in_win"), (win_num, "win_num")):
            _need(t, torch.int32, n)
        rc = _lib().dsvt_get_set_launch(
            ctypes.byref(self.p), _ptr(global_index), _ptr(coors_in_win), _ptr(voxel_num_in_win), _ptr(win_num),
            _ptr(self.global_index_in_set), _ptr(self.set_voxel_mask), _ptr(self.set_num),
            _ptr(self.mask_expand_0), _ptr(self.mask_expand_1), c_void_p(0), c_size_t(0), _stream())
        _check(rc, "dsvt_get_set_launch")
        return self


def gelu(x, voxel_num, out=None, zero_tails=1):
    """x [B,max_pillars,C] (or [max_pillars,C])."""
    _need(x, torch.float32, "x")
    _need(voxel_num, torch.int32, "voxel_num")
    B = x.shape[0] if x.dim() == 3 else 1
    out = torch.empty_like(x) if out is None else out
    p = GeluParams(B, x.shape[-2], x.shape[-1], zero_tails)
    _check(_lib().dsvt_gelu_launch(ctypes.byref(p), _ptr(x), _ptr(voxel_num), _ptr(out), _stream()), "dsvt_gelu_launch")
    return out


def layer_norm(x, voxel_num, gamma, beta, eps=0.0, residual=None, out=None, zero_tails=1):
    _need(x, torch.float32, "x")
    _need(voxel_num, torch.int32, "voxel_num")
    _need(gamma, torch.float32, "gamma")
    _need(beta, torch.float32, "beta")
    if residual is not None:
        _need(residual, torch.float32, "residual")
    B = x.shape[0] if x.dim() == 3 else 1
    out = torch.empty_like(x) if out is None else out
    p = LNParams(B, x.shape[-2], x.shape[-1], eps, zero_tails)
    _check(_lib().dsvt_layer_norm_launch(ctypes.byref(p), _ptr(x), _ptr(residual), _ptr(voxel_num), _ptr(gamma),
                                         _ptr(beta), _ptr(out), _stream()), "dsvt_layer_norm_launch")
    return out


class LnStage(Structure):
    _fields_ = [("residual", c_void_p), ("gamma", c_void_p), ("beta", c_void_p)]


def layer_norm_chain(x, voxel_num, stages, eps=0.0, out=None, zero_tails=1):
    """stages: list of (residual_or_None, gamma, beta); y = LN_n(...LN_1(x + r_1)... + r_n) in one launch."""
    _need(x, torch.float32, "x")
    B = x.shape[0] if x.dim() == 3 else 1
    out = torch.empty_like(x) if out is None else out
    arr = (LnStage * len(stages))()
    for i, (r, g, b) in enumerate(stages):
        arr[i] = LnStage(r.data_ptr() if r is not None else None, g.data_ptr(), b.data_ptr())
    p = LNParams(B, x.shape[-2], x.shape[-1], eps, zero_tails)
    _check(_lib().dsvt_layer_norm_chain_launch(ctypes.byref(p), _ptr(x), _ptr(voxel_num), arr, c_int32(len(stages)),
                                               _ptr(out), _stream()), "dsvt_layer_norm_chain_launch")
    return out


def filter_box(cfg, scores, classes, xs, ys, center, center_z, angle, dim, boxes=None, valid=None, zero_tails=1):
    B = scores.shape[0] if scores.dim() == 2 else 1
    K = cfg.max_top_k
    dev = scores.device
    boxes = torch.empty(B, K, 9, dtype=torch.float32, device=dev) if boxes is None else boxes
    valid = torch.empty(B, dtype=torch.int32, device=dev) if valid is None else valid
    p = FBParams(B, K, cfg.x_min, cfg.x_max, cfg.y_min, cfg.y_max, cfg.z_min, cfg.z_max,
                 cfg.voxel_x, cfg.voxel_y, cfg.voxel_z, cfg.score_threshold, zero_tails)
    _check(_lib().dsvt_filter_box_launch(ctypes.byref(p), _ptr(scores), _ptr(classes), _ptr(xs), _ptr(ys),
                                         _ptr(center), _ptr(center_z), _ptr(angle), _ptr(dim), _ptr(boxes),
                                         _ptr(valid), _stream()), "dsvt_filter_box_launch")
    return boxes, valid


class AttentionWeights:
    """Device-resident, pre-arranged weights of one attention layer (host float32 arrays in)."""

    def __init__(self, in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias, channels=192, heads=8):
        import numpy as np
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in
                (in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias)]
        self.channels, self.heads = channels, heads
        self.handle = _lib().dsvt_attention_weights_create(
            channels, heads, *[a.ctypes.data_as(c_void_p) for a in arrs])
        if not self.handle:
            raise DsvtError("dsvt_attention_weights_create: " + _lib().dsvt_last_error().decode())

    def close(self):
        if getattr(self, "handle", None):
            _lib().dsvt_attention_weights_destroy(c_void_p(self.handle))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def set_attention(weights, q, k, v, mask, set_num=None, out=None, precision=DSVT_ATTN_FP32, zero_tails=1, workspace=None):
    """Plugin-shaped form: q,k,v [B,max_sets,S,C] (or without B), mask [B,max_sets,H,S]."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (mask, "mask")):
        _need(t, torch.float32, n)
    B = q.shape[0] if q.dim() == 4 else 1
    max_sets, S, C = q.shape[-3], q.shape[-2], q.shape[-1]
    out = torch.empty_like(q) if out is None else out
    p = AttnParams(B, max_sets, S, C, weights.heads, 0, 0, precision, zero_tails)
    ws_bytes = int(_lib().dsvt_set_attention_workspace_size(ctypes.byref(p)))
    if ws_bytes and workspace is None:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
    rc = _lib().dsvt_set_attention_launch(ctypes.byref(p), c_void_p(weights.handle), _ptr(q), _ptr(k), _ptr(v),
                                          _ptr(mask), _ptr(set_num), _ptr(out), _ptr(workspace), c_size_t(ws_bytes), _stream())
    _check(rc, "dsvt_set_attention_launch")
    return out


def set_attention_workspace_bytes(B, max_sets, S, C, heads, max_pillars, precision):
    p = AttnParams(B, max_sets, S, C, heads, max_pillars, 0, precision, 1)
    return int(_lib().dsvt_set_attention_workspace_size(ctypes.byref(p)))


def set_attention_plan(global_index_in_set, mask, set_num, axis, max_pillars, heads=8, channels=192, out=None):
    """Builds the attention plan of one (window partition, axis) -- dsvt_set_attention_plan_launch.  The returned uint8
    tensor can be passed as `plan=` to every set_attention_fused call (GEMM-pipeline precisions) on that partition."""
    _need(global_index_in_set, torch.int32, "global_index_in_set")
    _need(mask, torch.float32, "mask")
    B = global_index_in_set.shape[0] if global_index_in_set.dim() == 4 else 1
    max_sets, S = global_index_in_set.shape[-2], global_index_in_set.shape[-1]
    p = AttnParams(B, max_sets, S, channels, heads, max_pillars, axis, DSVT_ATTN_FP32_TC, 1)
    lib = _lib()
    lib.dsvt_set_attention_plan_size.restype = c_size_t
    n = int(lib.dsvt_set_attention_plan_size(ctypes.byref(p)))
    out = torch.empty(n, dtype=torch.uint8, device=mask.device) if out is None else out
    rc = lib.dsvt_set_attention_plan_launch(ctypes.byref(p), _ptr(global_index_in_set), _ptr(mask), _ptr(set_num),
                                            _ptr(out), c_size_t(out.numel()), _stream())
    _check(rc, "dsvt_set_attention_plan_launch")
    return out


def set_attention_fused(weights, x, pos, global_index_in_set, mask, set_num, voxel_num, axis, out=None,
                        precision=DSVT_ATTN_FP32, zero_tails=1, workspace=None, plan=None, norm=None, stages=7, pos_table=None):
    """Fused gather + attention + scatter: x,pos [B,max_pillars,C], global_index_in_set [B,2,max_sets,S].
    `workspace`: uint8 device tensor of dsvt_set_attention_workspace_size bytes (allocated here when None and the
    precision needs one; pass a persistent buffer when capturing CUDA graphs).
    `norm` = (residual, gamma, beta, eps): out = LayerNorm(attention + residual) in the out-projection's epilogue
    (dsvt_set_attention_fused_norm_launch; GEMM-pipeline precisions).
    `stages` != 7: dsvt_set_attention_fused_stages_launch -- only the named kernels (1 QKV GEMM, 2 core, 4 out-projection).
    `pos_table` = (table [win_x * win_y, C], coors_in_win_2d [max_pillars, 3] int32, win_x): the position embedding as a table over
    the cells of a window (dsvt_set_attention_fused_table_launch; `pos` is then ignored and may be None)."""
    _need(x, torch.float32, "x")
    if pos_table is None:
        _need(pos, torch.float32, "pos")
    _need(global_index_in_set, torch.int32, "global_index_in_set")
    _need(mask, torch.float32, "mask")
    B = x.shape[0] if x.dim() == 3 else 1
    max_pillars, C = x.shape[-2], x.shape[-1]
    max_sets, S = global_index_in_set.shape[-2], global_index_in_set.shape[-1]
    out = torch.empty_like(x) if out is None else out
    p = AttnParams(B, max_sets, S, C, weights.heads, max_pillars, axis, precision, zero_tails)
    ws_bytes = int(_lib().dsvt_set_attention_workspace_size(ctypes.byref(p)))
    if ws_bytes and workspace is None:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    if pos_table is not None:
        tab, cells, win_x = pos_table
        _need(tab, torch.float32, "pos_table")
        _need(cells, torch.int32, "coors_in_win_2d")
        res, gamma, beta, eps = norm if norm is not None else (None, None, None, 0.0)
        rc = _lib().dsvt_set_attention_fused_table_launch(
            ctypes.byref(p), c_void_p(weights.handle), _ptr(x), _ptr(tab), _ptr(cells), c_int32(win_x), _ptr(global_index_in_set),
            _ptr(mask), _ptr(set_num), _ptr(voxel_num), _ptr(res), _ptr(gamma), _ptr(beta), c_float(eps), _ptr(out), _ptr(plan),
            _ptr(workspace), c_size_t(workspace.numel() if workspace is not None else 0), c_int32(stages), _stream())
        _check(rc, "dsvt_set_attention_fused_table_launch")
        return out
    if stages != 7:
        res, gamma, beta, eps = norm if norm is not None else (None, None, None, 0.0)
        rc = _lib().dsvt_set_attention_fused_stages_launch(
            ctypes.byref(p), c_void_p(weights.handle), _ptr(x), _ptr(pos), _ptr(global_index_in_set), _ptr(mask),
            _ptr(set_num), _ptr(voxel_num), _ptr(res), _ptr(gamma), _ptr(beta), c_float(eps), _ptr(out), _ptr(plan),
            _ptr(workspace), c_size_t(workspace.numel() if workspace is not None else 0), c_int32(stages), _stream())
        _check(rc, "dsvt_set_attention_fused_stages_launch")
        return out
    if norm is not None:
        res, gamma, beta, eps = norm
        for t, n in ((res, "residual"), (gamma, "gamma"), (beta, "beta")):
            _need(t, torch.float32, n)
        rc = _lib().dsvt_set_attention_fused_norm_launch(
            ctypes.byref(p), c_void_p(weights.handle), _ptr(x), _ptr(pos), _ptr(global_index_in_set), _ptr(mask),
            _ptr(set_num), _ptr(voxel_num), _ptr(res), _ptr(gamma), _ptr(beta), c_float(eps), _ptr(out), _ptr(plan),
            _ptr(workspace), c_size_t(workspace.numel() if workspace is not None else 0), _stream())
        _check(rc, "dsvt_set_attention_fused_norm_launch")
        return out
    if plan is not None and ws_bytes:
        rc = _lib().dsvt_set_attention_fused_planned_launch(
            ctypes.byref(p), c_void_p(weights.handle), _ptr(x), _ptr(pos), _ptr(global_index_in_set), _ptr(mask),
            _ptr(set_num), _ptr(voxel_num), _ptr(out), _ptr(plan), _ptr(workspace), c_size_t(workspace.numel()), _stream())
        _check(rc, "dsvt_set_attention_fused_planned_launch")
        return out
    rc = _lib().dsvt_set_attention_fused_launch(ctypes.byref(p), c_void_p(weights.handle), _ptr(x), _ptr(pos),
                                                _ptr(global_index_in_set), _ptr(mask), _ptr(set_num),
                                                _ptr(voxel_num), _ptr(out), _ptr(workspace),
                                                c_size_t(workspace.numel() if workspace is not None else 0), _stream())
    _check(rc, "dsvt_set_attention_fused_launch")
    return out


def get_value_by_index(x, pos, global_index_in_set, set_num, axis, zero_tails=1):
    B = x.shape[0] if x.dim() == 3 else 1
    max_pillars, C = x.shape[-2], x.shape[-1]
    max_sets, S = global_index_in_set.shape[-2], global_index_in_set.shape[-1]
    shape = (B, max_sets, S, C) if x.dim() == 3 else (max_sets, S, C)
    q, k, v = (torch.empty(shape, dtype=torch.float32, device=x.device) for _ in range(3))
    p = AttnParams(B, max_sets, S, C, 8, max_pillars, axis, 0, zero_tails)
    rc = _lib().dsvt_get_value_by_index_launch(ctypes.byref(p), _ptr(x), _ptr(pos), _ptr(global_index_in_set),
                                               _ptr(set_num), _ptr(q), _ptr(k), _ptr(v), _stream())
    _check(rc, "dsvt_get_value_by_index_launch")
    return q, k, v


def map_set_feature2voxel(feat, global_index_in_set, set_num, axis, max_pillars, zero_tails=1):
    B = feat.shape[0] if feat.dim() == 4 else 1
    max_sets, S, C = feat.shape[-3], feat.shape[-2], feat.shape[-1]
    shape = (B, max_pillars, C) if feat.dim() == 4 else (max_pillars, C)
    out = torch.empty(shape, dtype=torch.float32, device=feat.device)
    p = AttnParams(B, max_sets, S, C, 8, max_pillars, axis, 0, zero_tails)
    rc = _lib().dsvt_map_set_feature2voxel_launch(ctypes.byref(p), _ptr(feat), _ptr(global_index_in_set),
                                                  _ptr(set_num), _ptr(out), _stream())
    _check(rc, "dsvt_map_set_feature2voxel_launch")
    return out


class Linear:
    """y = x W^T + b on the tensor cores (tcgen05); W [N,K] float32 host array."""

    def __init__(self, W, b=None, precision=DSVT_LINEAR_TF32):
        import numpy as np
        W = np.ascontiguousarray(W, dtype=np.float32)
        self.N, self.K = W.shape
        bb = np.ascontiguousarray(b, dtype=np.float32) if b is not None else None
        self.handle = _lib().dsvt_linear_weights_create(self.N, self.K, W.ctypes.data_as(c_void_p),
                                                        bb.ctypes.data_as(c_void_p) if bb is not None else None,
                                                        precision)
        if not self.handle:
            raise DsvtError("dsvt_linear_weights_create: " + _lib().dsvt_last_error().decode())

    def __call__(self, x, out=None):
        _need(x, torch.float32, "x")
        M = x.shape[0]
        out = torch.empty(M, self.N, dtype=torch.float32, device=x.device) if out is None else out
        _check(_lib().dsvt_linear_launch(c_void_p(self.handle), _ptr(x), c_int32(M), _ptr(out), _stream()),
               "dsvt_linear_launch")
        return out

    def rows(self, x, rows, activation=0, out=None, zero_tails=1):
        """Plugin-style form (FP32_TC / FP16_GEMM weights): valid row count on the device, optional fused GELU
        (activation=1), rows [rows, max_rows) zero-filled.  x [max_rows, K] -> [max_rows, N]."""
        _need(x, torch.float32, "x")
        _need(rows, torch.int32, "rows")
        max_rows = x.shape[0]
        out = torch.empty(max_rows, self.N, dtype=torch.float32, device=x.device) if out is None else out
        _check(_lib().dsvt_linear_rows_launch(c_void_p(self.handle), _ptr(x), _ptr(rows), c_int32(max_rows),
                                              c_int32(activation), _ptr(out), c_int32(zero_tails), _stream()),
               "dsvt_linear_rows_launch")
        return out

    def rows_splitk(self, x, rows, add=None, out=None):
        """Split-K form (N == 192, K = 192 * kb): one launch, out[j] = partial product of K block j (bias and the optional
        residual rows `add` in part 0); the result is out.sum(0).  x [max_rows, K] -> out [kb, max_rows, 192]."""
        _need(x, torch.float32, "x")
        _need(rows, torch.int32, "rows")
        max_rows, kb = x.shape[0], self.K // 192
        out = torch.empty(kb, max_rows, self.N, dtype=torch.float32, device=x.device) if out is None else out
        _check(_lib().dsvt_linear_rows_splitk_launch(c_void_p(self.handle), _ptr(x), _ptr(add), _ptr(rows),
                                                     c_int32(max_rows), _ptr(out), _stream()),
               "dsvt_linear_rows_splitk_launch")
        return out

    def ffn_norm(self, second, x, rows, stages, eps=0.0, out=None, zero_tails=1):
        """self = Linear(192 -> 384), second = Linear(384 -> 192): gelu(x W1^T + b1) W2^T + b2 followed by the chain of
        (residual add + LayerNorm) stages, in ONE kernel (dsvt_ffn_fused_launch); x [max_rows, 192] -> [max_rows, 192]."""
        _need(x, torch.float32, "x")
        _need(rows, torch.int32, "rows")
        max_rows = x.shape[0]
        out = torch.empty(max_rows, second.N, dtype=torch.float32, device=x.device) if out is None else out
        arr = (LnStage * len(stages))()
        for i, (r, g, b) in enumerate(stages):
            arr[i] = LnStage(r.data_ptr() if r is not None else None, g.data_ptr(), b.data_ptr())
        _check(_lib().dsvt_ffn_fused_launch(c_void_p(self.handle), c_void_p(second.handle), _ptr(x), _ptr(rows), c_int32(max_rows),
                                            arr, c_int32(len(stages)), c_float(eps), _ptr(out), c_int32(zero_tails), _stream()),
               "dsvt_ffn_fused_launch")
        return out

    def rows_norm(self, x, rows, stages, eps=0.0, out=None, zero_tails=1):
        """Linear (N == 192, K in {192, 384}) + a chain of (residual add + LayerNorm) stages in one kernel:
        stages = [(residual_or_None, gamma, beta), ...] (<= 3).  x [max_rows, K] -> [max_rows, 192]."""
        _need(x, torch.float32, "x")
        _need(rows, torch.int32, "rows")
        max_rows = x.shape[0]
        out = torch.empty(max_rows, self.N, dtype=torch.float32, device=x.device) if out is None else out
        arr = (LnStage * len(stages))()
        for i, (r, g, b) in enumerate(stages):
            arr[i] = LnStage(r.data_ptr() if r is not None else None, g.data_ptr(), b.data_ptr())
        _check(_lib().dsvt_linear_rows_norm_launch(c_void_p(self.handle), _ptr(x), _ptr(rows), c_int32(max_rows), arr,
                                                   c_int32(len(stages)), c_float(eps), _ptr(out), c_int32(zero_tails),
                                                   _stream()), "dsvt_linear_rows_norm_launch")
        return out

    def rows_concat(self, x_lo, x_hi, rows, activation=0, out=None, zero_tails=1):
        """Same as rows() for the input [x_lo | x_hi] (two dense tensors, e.g. the PFN's concatenation of the point
        features and their per-pillar max), read in place.  activation: 0 none, 1 GELU, 2 ReLU."""
        _need(x_lo, torch.float32, "x_lo")
        _need(x_hi, torch.float32, "x_hi")
        _need(rows, torch.int32, "rows")
        max_rows, k_split = x_lo.shape[0], x_lo.shape[1]
        assert x_hi.shape[0] == max_rows and k_split + x_hi.shape[1] == self.K
        out = torch.empty(max_rows, self.N, dtype=torch.float32, device=x_lo.device) if out is None else out
        _check(_lib().dsvt_linear_rows_concat_launch(c_void_p(self.handle), _ptr(x_lo), _ptr(x_hi), c_int32(k_split),
                                                     _ptr(rows), c_int32(max_rows), c_int32(activation), _ptr(out),
                                                     c_int32(zero_tails), _stream()),
               "dsvt_linear_rows_concat_launch")
        return out

    def close(self):
        if getattr(self, "handle", None):
            _lib().dsvt_linear_weights_destroy(c_void_p(self.handle))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SmallLinear:
    """y = act((x W^T) * scale + shift) for the narrow first layers of the PFN / position-embedding MLPs (K = 2, 4, 10);
    scale / shift = the folded BatchNorm1d.  W [N,K] float32 host array."""

    def __init__(self, W, scale=None, shift=None):
        import numpy as np
        W = np.ascontiguousarray(W, dtype=np.float32)
        self.N, self.K = W.shape
        sc = np.ascontiguousarray(scale, dtype=np.float32) if scale is not None else None
        sh = np.ascontiguousarray(shift, dtype=np.float32) if shift is not None else None
        self.handle = _lib().dsvt_small_linear_create(self.N, self.K, W.ctypes.data_as(c_void_p),
                                                      sc.ctypes.data_as(c_void_p) if sc is not None else None,
                                                      sh.ctypes.data_as(c_void_p) if sh is not None else None)
        if not self.handle:
            raise DsvtError("dsvt_small_linear_create: " + _lib().dsvt_last_error().decode())

    def __call__(self, x, rows, activation=2, out=None, zero_tails=1):
        """x [max_rows, K] (or [B, max_rows, K]) -> [.., max_rows, N]; rows [B] int32 on the device."""
        _need(x, torch.float32, "x")
        _need(rows, torch.int32, "rows")
        B = x.shape[0] if x.dim() == 3 else 1
        max_rows = x.shape[-2]
        assert x.shape[-1] == self.K
        shape = (B, max_rows, self.N) if x.dim() == 3 else (max_rows, self.N)
        out = torch.empty(shape, dtype=torch.float32, device=x.device) if out is None else out
        _check(_lib().dsvt_small_linear_launch(c_void_p(self.handle), _ptr(x), _ptr(rows), c_int32(B), c_int32(max_rows),
                                               c_int32(activation), _ptr(out), c_int32(zero_tails), _stream()),
               "dsvt_small_linear_launch")
        return out

    def close(self):
        if getattr(self, "handle", None):
            _lib().dsvt_small_linear_destroy(c_void_p(self.handle))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pos_embed_mlp(first, second, x2, rows, out=None, zero_tails=1):
    """Position-embedding MLP Linear(2->192)+BN+ReLU -> Linear(192->192) in one kernel (dsvt_pos_embed_mlp_launch):
    first = SmallLinear(2 -> 192), second = Linear(192 -> 192, FP32_TC / FP16_GEMM); x2 [max_rows, 2]."""
    _need(x2, torch.float32, "x2")
    _need(rows, torch.int32, "rows")
    max_rows = x2.shape[-2]
    out = torch.empty(max_rows, 192, dtype=torch.float32, device=x2.device) if out is None else out
    _check(_lib().dsvt_pos_embed_mlp_launch(c_void_p(first.handle), c_void_p(second.handle), _ptr(x2), _ptr(rows),
                                            c_int32(max_rows), _ptr(out), c_int32(zero_tails), _stream()),
           "dsvt_pos_embed_mlp_launch")
    return out


def pos_embed_mlp_batch(firsts, seconds, x2s, rows, outs, zero_tails=1):
    """n (<= 8) position-embedding MLPs in one launch (dsvt_pos_embed_mlp_batch_launch): lists of SmallLinear(2 -> 192),
    Linear(192 -> 192), coordinate tensors [max_rows, 2] and output tensors [max_rows, 192]."""
    n = len(firsts)
    assert n == len(seconds) == len(x2s) == len(outs) and 1 <= n <= 8
    for t in x2s:
        _need(t, torch.float32, "x2")
    _need(rows, torch.int32, "rows")
    arr = lambda vals: (c_void_p * n)(*vals)
    _check(_lib().dsvt_pos_embed_mlp_batch_launch(arr([f.handle for f in firsts]), arr([s_.handle for s_ in seconds]),
                                                  arr([t.data_ptr() for t in x2s]), c_int32(n), _ptr(rows),
                                                  c_int32(x2s[0].shape[-2]), arr([o.data_ptr() for o in outs]),
                                                  c_int32(zero_tails), _stream()), "dsvt_pos_embed_mlp_batch_launch")
    return outs


def attention_tail_ffn(weights, fc1, fc2, x, global_index_in_set, voxel_num, axis, plan, workspace, norm1, stages, eps=0.0,
                       src=None, out=None, zero_tails=1):
    """The tail of an encoder layer in one kernel (dsvt_attention_tail_ffn_launch): out-projection of the attention whose QKV
    projection + core have just run (set_attention_fused(..., stages=3) with the same weights / plan / workspace),
    norm1 = (gamma, beta, eps) of LayerNorm(attention + x), the FFN fc1 -> GELU -> fc2 and the LayerNorm chain `stages`
    (the first stage's residual is the `src` this call writes)."""
    _need(x, torch.float32, "x")
    max_pillars, C = x.shape[-2], x.shape[-1]
    max_sets, S = global_index_in_set.shape[-2], global_index_in_set.shape[-1]
    p = AttnParams(1, max_sets, S, C, weights.heads, max_pillars, axis, DSVT_ATTN_FP32_TC, zero_tails)
    src = torch.empty_like(x) if src is None else src
    out = torch.empty_like(x) if out is None else out
    arr = (LnStage * len(stages))()
    for i, (r, g, b) in enumerate(stages):
        arr[i] = LnStage(r.data_ptr() if r is not None else None, g.data_ptr(), b.data_ptr())
    g1, b1, e1 = norm1
    _check(_lib().dsvt_attention_tail_ffn_launch(ctypes.byref(p), c_void_p(weights.handle), _ptr(plan), _ptr(workspace),
                                                 c_size_t(workspace.numel()), _ptr(voxel_num), _ptr(x), _ptr(g1), _ptr(b1),
                                                 c_float(e1), c_void_p(fc1.handle), c_void_p(fc2.handle), arr,
                                                 c_int32(len(stages)), c_float(eps), _ptr(src), _ptr(out), _stream()),
           "dsvt_attention_tail_ffn_launch")
    return out


def vfe_fused(pfn0, pfn1, point_features, point_index_in_voxel, voxel_num, point_num, out=None, workspace=None, zero_tails=1):
    """The pillar feature net in one kernel (dsvt_vfe_fused_launch): pfn0 = SmallLinear(10 -> 96), pfn1 = Linear(192 -> 192,
    FP32_TC); point_features [max_points, 10], point_index_in_voxel [max_pillars, npv] -> voxel features [max_pillars, 192]."""
    _need(point_features, torch.float32, "point_features")
    _need(point_index_in_voxel, torch.int32, "point_index_in_voxel")
    max_points = point_features.shape[-2]
    max_pillars, npv = point_index_in_voxel.shape[-2], point_index_in_voxel.shape[-1]
    out = torch.empty(max_pillars, 192, dtype=torch.float32, device=point_features.device) if out is None else out
    ws = int(_lib().dsvt_vfe_fused_workspace_size(c_int32(max_points), c_int32(npv)))
    if workspace is None:
        workspace = torch.empty(ws, dtype=torch.uint8, device=point_features.device)
    _check(_lib().dsvt_vfe_fused_launch(c_void_p(pfn0.handle), c_void_p(pfn1.handle), _ptr(point_features),
                                        _ptr(point_index_in_voxel), _ptr(voxel_num), _ptr(point_num), c_int32(max_points),
                                        c_int32(max_pillars), c_int32(npv), _ptr(out), _ptr(workspace),
                                        c_size_t(workspace.numel()), c_int32(zero_tails), _stream()), "dsvt_vfe_fused_launch")
    return out


class ScatterMaxParams(Structure):
    _fields_ = [("batch", c_int32), ("max_points_num", c_int32), ("max_pillars_num", c_int32), ("feature_num", c_int32),
                ("max_num_points_per_voxel", c_int32), ("zero_tails", c_int32)]


class Map2BevParams(Structure):
    _fields_ = [("batch", c_int32), ("max_pillars_num", c_int32), ("channel_num", c_int32), ("grid_size_x", c_int32),
                ("grid_size_y", c_int32)]


def torch_scatter_max(point_features, point_index_in_voxel, point_num_in_voxel, voxel_num, point_num=None,
                      max_point=None, max_voxel=None, zero_tails=1):
    """TorchScatterMaxPlugin: per-pillar channel-wise max, broadcast back to the pillar's point rows."""
    _need(point_features, torch.float32, "point_features")
    _need(point_index_in_voxel, torch.int32, "point_index_in_voxel")
    B = point_features.shape[0] if point_features.dim() == 3 else 1
    max_points, F = point_features.shape[-2], point_features.shape[-1]
    max_pillars, npv = point_index_in_voxel.shape[-2], point_index_in_voxel.shape[-1]
    max_point = torch.empty_like(point_features) if max_point is None else max_point
    vshape = (B, max_pillars, F) if point_features.dim() == 3 else (max_pillars, F)
    max_voxel = torch.empty(vshape, dtype=torch.float32, device=point_features.device) if max_voxel is None else max_voxel
    p = ScatterMaxParams(B, max_points, max_pillars, F, npv, zero_tails)
    rc = _lib().dsvt_torch_scatter_max_launch(ctypes.byref(p), _ptr(point_features), _ptr(point_index_in_voxel),
                                              _ptr(point_num_in_voxel), _ptr(voxel_num), _ptr(point_num),
                                              _ptr(max_point), _ptr(max_voxel), _stream())
    _check(rc, "dsvt_torch_scatter_max_launch")
    return max_point, max_voxel


def map2bev(voxel_features, coords, voxel_num, grid_x, grid_y, out=None):
    """Map2BevPlugin: dense [grid_y, grid_x, C] BEV map from the pillar rows."""
    _need(voxel_features, torch.float32, "voxel_features")
    _need(coords, torch.int32, "coords")
    B = voxel_features.shape[0] if voxel_features.dim() == 3 else 1
    max_pillars, C = voxel_features.shape[-2], voxel_features.shape[-1]
    shape = (B, grid_y, grid_x, C) if voxel_features.dim() == 3 else (grid_y, grid_x, C)
    out = torch.empty(shape, dtype=torch.float32, device=voxel_features.device) if out is None else out
    p = Map2BevParams(B, max_pillars, C, grid_x, grid_y)
    rc = _lib().dsvt_map2bev_launch(ctypes.byref(p), _ptr(voxel_features), _ptr(coords), _ptr(voxel_num), _ptr(out), _stream())
    _check(rc, "dsvt_map2bev_launch")
    return out


class CenterHeadParams(Structure):
    _fields_ = [("batch", c_int32), ("num_classes", c_int32), ("height", c_int32), ("width", c_int32), ("max_top_k", c_int32)]


class NmsParams(Structure):
    _fields_ = [("batch", c_int32), ("max_boxes", c_int32), ("nms_thresh", c_float), ("zero_tails", c_int32)]


class CenterHeadTopK:
    """The CenterHead post-process graph in front of FilterBoxByScorePlugin (dsvt_center_head_topk_launch): pre-allocated
    outputs in the plugin's input order -- scores, classes, xs, ys, center, center_z, angle, dim."""

    def __init__(self, num_classes, height, width, max_top_k=500, batch=1, device="cuda"):
        self.p = CenterHeadParams(batch, num_classes, height, width, max_top_k)
        lib = _lib()
        lib.dsvt_center_head_topk_workspace_size.restype = c_size_t
        n = int(lib.dsvt_center_head_topk_workspace_size(ctypes.byref(self.p)))
        if n == 0:
            raise DsvtError("center_head_topk: " + lib.dsvt_last_error().decode())
        B, K = batch, max_top_k
        self.ws = torch.empty(n, dtype=torch.uint8, device=device)
        self.scores = torch.empty(B, K, dtype=torch.float32, device=device)
        self.classes, self.xs, self.ys = (torch.empty(B, K, dtype=torch.int32, device=device) for _ in range(3))
        self.center = torch.empty(B, 1, K, 2, dtype=torch.float32, device=device)
        self.center_z = torch.empty(B, 1, K, 1, dtype=torch.float32, device=device)
        self.angle = torch.empty(B, 1, K, 1, dtype=torch.float32, device=device)
        self.dim = torch.empty(B, 1, K, 3, dtype=torch.float32, device=device)

    @property
    def outputs(self):
        return [self.scores, self.classes, self.xs, self.ys, self.center, self.center_z, self.angle, self.dim]

    def __call__(self, heatmap, center, center_z, dim, rot):
        for t, n in ((heatmap, "heatmap"), (center, "center"), (center_z, "center_z"), (dim, "dim"), (rot, "rot")):
            _need(t, torch.float32, n)
        rc = _lib().dsvt_center_head_topk_launch(
            ctypes.byref(self.p), _ptr(heatmap), _ptr(center), _ptr(center_z), _ptr(dim), _ptr(rot), _ptr(self.scores),
            _ptr(self.classes), _ptr(self.xs), _ptr(self.ys), _ptr(self.center), _ptr(self.center_z), _ptr(self.angle),
            _ptr(self.dim), _ptr(self.ws), c_size_t(self.ws.numel()), _stream())
        _check(rc, "dsvt_center_head_topk_launch")
        return self


class RotatedNms:
    """nms_cpu of the reference (include/helper.h:257-283) on the GPU: dsvt_rotated_nms_launch."""

    def __init__(self, max_boxes=500, nms_thresh=0.01, batch=1, device="cuda", zero_tails=1):
        self.p = NmsParams(batch, max_boxes, nms_thresh, zero_tails)
        lib = _lib()
        lib.dsvt_rotated_nms_workspace_size.restype = c_size_t
        n = int(lib.dsvt_rotated_nms_workspace_size(ctypes.byref(self.p)))
        if n == 0:
            raise DsvtError("rotated_nms: " + lib.dsvt_last_error().decode())
        self.ws = torch.empty(n, dtype=torch.uint8, device=device)
        self.boxes = torch.empty(batch, max_boxes, 9, dtype=torch.float32, device=device)
        self.num = torch.empty(batch, dtype=torch.int32, device=device)
        self.keep = torch.empty(batch, max_boxes, dtype=torch.int32, device=device)

    def __call__(self, boxes, valid):
        _need(boxes, torch.float32, "boxes")
        _need(valid, torch.int32, "valid")
        rc = _lib().dsvt_rotated_nms_launch(ctypes.byref(self.p), _ptr(boxes), _ptr(valid), _ptr(self.boxes), _ptr(self.num),
                                            _ptr(self.keep), _ptr(self.ws), c_size_t(self.ws.numel()), _stream())
        _check(rc, "dsvt_rotated_nms_launch")
        return self
