"""Host-side mirror of the reference's plugin interface, for tests and tools.

``PluginLibrary`` opens a shared object that exports the C harness of
``include/dsvt_b200_plugin_c.h`` -- our ``libdsvt_b200_plugins.so`` or one of the
``oracle/_ref/libref_<plugin>.so`` built from the reference's unmodified sources -- and drives a
plugin exactly the way TensorRT does: creator lookup by name/version, ``createPlugin`` with a
``PluginFieldCollection``, ``getOutputDimensions``, ``getWorkspaceSize``, ``enqueue``, ``serialize``.

The ``add_*`` helpers carry the names and argument order of the reference's
``include/plugin_helper.h`` (``add_voxel_generator`` :15, ``add_window_partition`` :174,
``add_get_set_op`` :253, ``add_layer_norm_op`` :489, ``add_gelu_op`` :557,
``add_filter_box_by_score_op`` :607) and, like them, forward only the fields the creator advertises.
"""
import ctypes
from ctypes import POINTER, Structure, c_char_p, c_int32, c_size_t, c_void_p

import numpy as np
import torch

from ._lib import load_plugin_library

FIELD_FLOAT32, FIELD_INT32 = 1, 5
DTYPE_FLOAT, DTYPE_INT32 = 0, 3
_TORCH = {DTYPE_FLOAT: torch.float32, DTYPE_INT32: torch.int32}
_FROM_TORCH = {torch.float32: DTYPE_FLOAT, torch.int32: DTYPE_INT32}


class _Field(Structure):
    _fields_ = [("name", c_char_p), ("data", c_void_p), ("type", c_int32), ("length", c_int32)]


class _Desc(Structure):
    _fields_ = [("nb_dims", c_int32), ("dims", c_int32 * 8), ("dtype", c_int32)]


def _desc_of(t):
    d = _Desc()
    d.nb_dims = t.dim()
    for i, s in enumerate(t.shape):
        d.dims[i] = s
    d.dtype = _FROM_TORCH[t.dtype]
    return d


class PluginLibrary:
    def __init__(self, path=None):
        if path is None:
            self.lib = load_plugin_library()
        else:
            # a foreign harness library (oracle/_ref/libref_*.so: the reference's own plugin sources): opened on its own --
            # it must not pull this repo's kernels into the process (the reference bench arm runs none of them)
            self.lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        L = self.lib
        L.dsvt_plugin_registry_name.restype = c_char_p
        L.dsvt_plugin_create.restype = c_void_p
        L.dsvt_plugin_deserialize.restype = c_void_p
        L.dsvt_plugin_clone.restype = c_void_p
        L.dsvt_plugin_type.restype = c_char_p
        L.dsvt_plugin_version.restype = c_char_p
        L.dsvt_plugin_serialization_size.restype = c_size_t
        L.dsvt_plugin_workspace_size.restype = c_size_t
        for f in ("dsvt_plugin_clone", "dsvt_plugin_destroy", "dsvt_plugin_type", "dsvt_plugin_version",
                  "dsvt_plugin_nb_outputs", "dsvt_plugin_serialization_size"):
            getattr(L, f).argtypes = [c_void_p]
        L.dsvt_plugin_serialize.argtypes = [c_void_p, c_void_p]

    def registered(self):
        n = self.lib.dsvt_plugin_registry_size()
        return [self.lib.dsvt_plugin_registry_name(i).decode() for i in range(n)]

    def field_names(self, name, version="1"):
        arr = (c_char_p * 32)()
        n = self.lib.dsvt_plugin_field_names(name.encode(), version.encode(), arr, 32)
        if n < 0:
            raise KeyError(f"no creator registered for {name} v{version}")
        return [arr[i].decode() for i in range(n)]

    def create(self, name, fields, version="1", filter_advertised=True):
        """fields: dict name -> int | float | sequence | numpy array (int32 / float32)."""
        keep = []
        advertised = self.field_names(name, version)
        arr = (_Field * len(fields))()
        n = 0
        for key, val in fields.items():
            if filter_advertised and key not in advertised:
                continue   # plugin_helper.h only emits names returned by getFieldNames()
            a = np.asarray(val)
            if a.dtype.kind in "iub":
                a = np.ascontiguousarray(a.reshape(-1), dtype=np.int32)
                typ = FIELD_INT32
            else:
                a = np.ascontiguousarray(a.reshape(-1), dtype=np.float32)
                typ = FIELD_FLOAT32
            keep.append(a)
            arr[n] = _Field(key.encode(), a.ctypes.data_as(c_void_p), typ, a.size)
            n += 1
        h = self.lib.dsvt_plugin_create(name.encode(), version.encode(), arr, n)
        if not h:
            raise RuntimeError(f"createPlugin({name}) failed")
        return Plugin(self, h)

    def deserialize(self, name, blob, version="1"):
        buf = ctypes.create_string_buffer(bytes(blob), len(blob))
        h = self.lib.dsvt_plugin_deserialize(name.encode(), version.encode(), buf, c_size_t(len(blob)))
        if not h:
            raise RuntimeError(f"deserializePlugin({name}) failed")
        return Plugin(self, h)


class Plugin:
    def __init__(self, library, handle):
        self.library, self.lib, self.h = library, library.lib, handle
        self._ws = None

    @property
    def type(self):
        return self.lib.dsvt_plugin_type(self.h).decode()

    @property
    def version(self):
        return self.lib.dsvt_plugin_version(self.h).decode()

    @property
    def nb_outputs(self):
        return self.lib.dsvt_plugin_nb_outputs(self.h)

    def serialize(self):
        n = self.lib.dsvt_plugin_serialization_size(self.h)
        buf = ctypes.create_string_buffer(n)
        self.lib.dsvt_plugin_serialize(self.h, buf)
        return buf.raw

    def clone(self):
        h = self.lib.dsvt_plugin_clone(self.h)
        if not h:
            raise RuntimeError("clone failed")
        return Plugin(self.library, h)

    def output_descs(self, in_descs):
        arr = (_Desc * len(in_descs))(*in_descs)
        outs = []
        for i in range(self.nb_outputs):
            o = _Desc()
            rc = self.lib.dsvt_plugin_output_desc(c_void_p(self.h), i, arr, len(in_descs), ctypes.byref(o))
            if rc != 0:
                raise RuntimeError("getOutputDimensions failed")
            outs.append(o)
        return outs

    def supports_format(self, pos, descs, nb_inputs):
        arr = (_Desc * len(descs))(*descs)
        return bool(self.lib.dsvt_plugin_supports_format(c_void_p(self.h), pos, arr, nb_inputs, len(descs) - nb_inputs))

    def enqueue(self, inputs, outputs=None, poison=None):
        """inputs: list of CUDA tensors.  Allocates outputs (optionally pre-filled with `poison`) and workspace."""
        in_descs = [_desc_of(t) for t in inputs]
        out_descs = self.output_descs(in_descs)
        dev = inputs[0].device
        if outputs is None:
            outputs = []
            for o in out_descs:
                shape = [o.dims[i] for i in range(o.nb_dims)]
                t = torch.empty(shape, dtype=_TORCH[o.dtype], device=dev)
                if poison is not None:
                    t.fill_(poison if t.dtype.is_floating_point else -7)
                outputs.append(t)
        ia = (_Desc * len(in_descs))(*in_descs)
        oa = (_Desc * len(out_descs))(*out_descs)
        ws_bytes = int(self.lib.dsvt_plugin_workspace_size(c_void_p(self.h), ia, len(in_descs), oa, len(out_descs)))
        if self._ws is None or self._ws.numel() < ws_bytes or self._ws.device != dev:
            self._ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        ip = (c_void_p * len(inputs))(*[t.data_ptr() for t in inputs])
        op = (c_void_p * len(outputs))(*[t.data_ptr() for t in outputs])
        rc = self.lib.dsvt_plugin_enqueue(c_void_p(self.h), ia, len(inputs), oa, len(outputs), ip, op,
                                          c_void_p(self._ws.data_ptr()),
                                          c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError(f"{self.type}.enqueue returned {rc}")
        return outputs

    def destroy(self):
        if self.h:
            self.lib.dsvt_plugin_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


# ---- mirrors of include/plugin_helper.h --------------------------------------------------------
def add_voxel_generator(lib, max_points_num, max_points_num_voxel_filter, max_pillars_num, point_feature_num,
                        feature_num, max_num_points_per_voxel, x_min, x_max, y_min, y_max, z_min, z_max,
                        voxel_size_x, voxel_size_y, voxel_size_z, grid_size_x, grid_size_y, grid_size_z):
    return lib.create("Points2FeaturesPlugin", {
        "max_points_num": max_points_num, "max_points_num_voxel_filter": max_points_num_voxel_filter,
        "max_pillars_num": max_pillars_num, "point_feature_num": point_feature_num, "feature_num": feature_num,
        "max_num_points_per_voxel": max_num_points_per_voxel,
        "point_cloud_range": [x_min, y_min, z_min, x_max, y_max, z_max],      # plugin_helper.h:32-37
        "voxel_size": [voxel_size_x, voxel_size_y, voxel_size_z],
        "grid_size": [grid_size_x, grid_size_y, grid_size_z]})


def add_window_partition(lib, max_win_num, max_voxel_num_per_win, sparse_shape, win_shape, shift_list):
    return lib.create("WindowPartitionPlugin", {
        "max_win_num": max_win_num, "max_voxel_num_per_win": max_voxel_num_per_win,
        "sparse_shape": list(sparse_shape), "win_shape": list(win_shape), "shift_list": list(shift_list)})


def add_get_set_op(lib, max_win_num, max_voxel_num_per_win, voxel_num_set, win_shape):
    return lib.create("GetSetPlugin", {
        "max_win_num": max_win_num, "max_voxel_num_per_win": max_voxel_num_per_win,
        "voxel_num_set": voxel_num_set, "win_shape": list(win_shape)})


def add_gelu_op(lib, max_pillars_num, channel_num):
    return lib.create("GeluPlugin", {"max_pillars_num": max_pillars_num, "channel_num": channel_num})


def add_layer_norm_op(lib, max_pillars_num, channel_num, weights, bias, eps=1e-5):
    # the helper offers "eps" but the creator advertises "pes", so eps is never forwarded
    # (plugin_helper.h:527 vs layerNorm.cu:497) -- reproduced by filter_advertised
    w = np.asarray(weights, dtype=np.float32)
    return lib.create("LayerNormPlugin", {
        "max_pillars_num": max_pillars_num, "channel_num": channel_num, "weights_size": int(w.size),
        "eps": float(eps), "weights": w, "bias": np.asarray(bias, dtype=np.float32)})


def add_filter_box_by_score_op(lib, max_top_k, x_min, x_max, y_min, y_max, z_min, z_max, voxel_x, voxel_y, voxel_z,
                               score_threshold):
    return lib.create("FilterBoxByScorePlugin", {
        "max_top_k": max_top_k,
        "point_cloud_range": [x_min, x_max, y_min, y_max, z_min, z_max],      # plugin_helper.h:627-632
        "voxel_size": [voxel_x, voxel_y, voxel_z], "score_threshold": float(score_threshold)})


def add_get_value_by_index_op(lib, max_win_num, voxel_num_set, channel_num, axis_id):
    return lib.create("GetValueByIndexPlugin", {
        "max_win_num": max_win_num, "voxel_num_set": voxel_num_set, "channel_num": channel_num, "axis_id": axis_id})


def add_map_set_feature2voxel_op(lib, max_win_num, voxel_num_set, channel_num, axis_id, max_pillars_num):
    return lib.create("MapSetFeature2VoxelPlugin", {
        "max_win_num": max_win_num, "voxel_num_set": voxel_num_set, "channel_num": channel_num,
        "axis_id": axis_id, "max_pillars_num": max_pillars_num})


def add_set_attention_op(lib, max_win_num, voxel_num_set, channel_num, num_heads, in_proj_weight, in_proj_bias,
                         out_proj_weight, out_proj_bias, precision=0):
    """New plugin replacing the multHeadAttention() layer sub-graph (src/dsvt-ai-trt.cpp:288-458)."""
    return lib.create("SetAttentionPlugin", {
        "max_win_num": max_win_num, "voxel_num_set": voxel_num_set, "channel_num": channel_num,
        "num_heads": num_heads, "precision": precision,
        "in_proj_weight": np.asarray(in_proj_weight, np.float32), "in_proj_bias": np.asarray(in_proj_bias, np.float32),
        "out_proj_weight": np.asarray(out_proj_weight, np.float32),
        "out_proj_bias": np.asarray(out_proj_bias, np.float32)})


def add_set_attention_fused_op(lib, max_win_num, voxel_num_set, channel_num, num_heads, max_pillars_num, axis_id,
                               in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias, precision=3,
                               norm_weight=None, norm_bias=None, norm_eps=0.0):
    """One node for GetValueByIndexPlugin -> multHeadAttention() -> MapSetFeature2VoxelPlugin
    (src/dsvt-ai-trt.cpp:653-668).  precision: DSVT_ATTN_* (3 = FP32-accurate tcgen05 GEMM pipeline).
    norm_weight / norm_bias: also folds the kSUM + LayerNormPlugin behind it (norm1(y + x), :669-676) into the node."""
    fields = {
        "max_win_num": max_win_num, "voxel_num_set": voxel_num_set, "channel_num": channel_num,
        "num_heads": num_heads, "precision": precision, "max_pillars_num": max_pillars_num, "axis_id": axis_id,
        "in_proj_weight": np.asarray(in_proj_weight, np.float32), "in_proj_bias": np.asarray(in_proj_bias, np.float32),
        "out_proj_weight": np.asarray(out_proj_weight, np.float32),
        "out_proj_bias": np.asarray(out_proj_bias, np.float32)}
    if norm_weight is not None:
        fields.update({"norm_weight": np.asarray(norm_weight, np.float32), "norm_bias": np.asarray(norm_bias, np.float32),
                       "norm_eps": float(norm_eps)})
    return lib.create("SetAttentionFusedPlugin", fields)


def add_layer_norm_chain_op(lib, max_pillars_num, channel_num, weights, bias, eps=0.0):
    """n = len(weights) consecutive (kSUM + LayerNormPlugin) pairs as one node; weights / bias [n, C].
    Inputs: x, voxel_num, residual_1 .. residual_n."""
    w, b = np.asarray(weights, np.float32), np.asarray(bias, np.float32)
    return lib.create("LayerNormChainPlugin", {
        "max_pillars_num": max_pillars_num, "channel_num": channel_num, "n_stages": int(w.shape[0]), "eps": float(eps),
        "weights": w, "bias": b})


def add_linear_op(lib, max_rows, in_features, out_features, weight, bias, activation=0, precision=3, ln_weights=None,
                  ln_bias=None, ln_eps=0.0):
    """A TensorRT FullyConnected layer of the 3-D backbone (+ GELU / ReLU, or + the LayerNorm chain that follows the FFN) as
    one node.  Inputs: x [B,max_rows,K], rows [B], residual_1 .. residual_n (n = len(ln_weights))."""
    fields = {"max_rows": max_rows, "in_features": in_features, "out_features": out_features, "activation": activation,
              "precision": precision, "weight": np.asarray(weight, np.float32), "bias": np.asarray(bias, np.float32),
              "n_stages": 0 if ln_weights is None else int(np.asarray(ln_weights).shape[0]), "ln_eps": float(ln_eps)}
    if ln_weights is not None:
        fields.update({"ln_weights": np.asarray(ln_weights, np.float32), "ln_bias": np.asarray(ln_bias, np.float32)})
    return lib.create("LinearPlugin", fields)


def add_ffn_fused_op(lib, max_rows, weight1, bias1, weight2, bias2, ln_weights, ln_bias, ln_eps=0.0):
    """fullyConnected_gelu_fullyConnected (src/dsvt-ai-trt.cpp:494-529) + the n = len(ln_weights) (kSUM + LayerNormPlugin) pairs
    behind it as one node / one kernel.  Inputs: x [B,max_rows,192], rows [B], residual_1 .. residual_n."""
    g = np.asarray(ln_weights, np.float32)
    return lib.create("FfnFusedPlugin", {
        "max_rows": max_rows, "weight1": np.asarray(weight1, np.float32), "bias1": np.asarray(bias1, np.float32),
        "weight2": np.asarray(weight2, np.float32), "bias2": np.asarray(bias2, np.float32), "n_stages": int(g.shape[0]),
        "ln_eps": float(ln_eps), "ln_weights": g, "ln_bias": np.asarray(ln_bias, np.float32)})


def add_vfe_fused_op(lib, max_points_num, max_pillars_num, max_num_points_per_voxel, weight0, scale0, shift0, weight1, bias1):
    """The pillar feature net of src/dsvt-ai-trt.cpp:571-590 as one node: PFN 0 (Linear(10 -> 96), BatchNorm as scale0 / shift0,
    ReLU), scatter-max, concat, PFN 1 (Linear(192 -> 192) with its BatchNorm folded into weight1 / bias1, ReLU), scatter-max.
    Inputs: Points2FeaturesPlugin outputs 0, 1, 4, 5."""
    return lib.create("VfeFusedPlugin", {
        "max_points_num": max_points_num, "max_pillars_num": max_pillars_num,
        "max_num_points_per_voxel": max_num_points_per_voxel, "weight0": np.asarray(weight0, np.float32),
        "scale0": np.asarray(scale0, np.float32), "shift0": np.asarray(shift0, np.float32),
        "weight1": np.asarray(weight1, np.float32), "bias1": np.asarray(bias1, np.float32)})


def add_torch_scatter_max(lib, max_points_num, max_pillars_num, feature_num):
    """include/plugin_helper.h:125 (inputs: point_features, point_index_in_voxel, point_num_in_voxel, voxel_num)."""
    return lib.create("TorchScatterMaxPlugin", {
        "max_points_num": max_points_num, "max_pillars_num": max_pillars_num, "feature_num": feature_num})


def add_map_2_bev_op(lib, max_pillars_num, channel_num, grid_size_x, grid_size_y):
    """include/plugin_helper.h:371 (inputs: voxel_features, coords, valid_voxel_num)."""
    return lib.create("Map2BevPlugin", {
        "max_pillars_num": max_pillars_num, "channel_num": channel_num, "grid_size_x": grid_size_x,
        "grid_size_y": grid_size_y})


def add_set_attention_plan_op(lib, max_win_num, voxel_num_set, num_heads, max_pillars_num, axis_id):
    """Token-order plan of one (window partition, axis): inputs global_index_in_set, mask, set_num (GetSetPlugin outputs
    0, 3, 2); its output is the optional 7th input of SetAttentionFusedPlugin."""
    return lib.create("SetAttentionPlanPlugin", {
        "max_win_num": max_win_num, "voxel_num_set": voxel_num_set, "num_heads": num_heads,
        "max_pillars_num": max_pillars_num, "axis_id": axis_id})
