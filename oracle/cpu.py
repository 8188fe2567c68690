"""numpy front-end of the C oracle (oracle/dsvt_oracle.c).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build_oracle()
        _lib = ctypes.CDLL(path)
        _lib.oracle_points2features.restype = ctypes.c_int
        _lib.oracle_window_partition.restype = ctypes.c_int
        _lib.oracle_get_set.restype = ctypes.c_int
        _lib.oracle_set_attention.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


c_int, c_float = ctypes.c_int, ctypes.c_float


def points2features(points, n_points, cfg):
    """points [>=n,4] f32 -> dict of the six plugin outputs (canonical order)."""
    points = _f32(points)
    mr, mp, npv = cfg.max_points_num_voxel_filter, cfg.max_pillars_num, cfg.max_num_points_per_voxel
    out = dict(point_features=np.zeros((mr, 10), np.float32), point_index_in_voxel=np.zeros((mp, npv), np.int32),
               coords=np.zeros((mp, 4), np.int32), point_num_in_voxel=np.zeros((mp,), np.int32))
    pn, ptn = c_int(0), c_int(0)
    rc = lib().oracle_points2features(
        _p(points), c_int(int(n_points)), c_int(cfg.max_points_num), c_int(mr), c_int(mp), c_int(npv),
        c_float(cfg.x_min), c_float(cfg.x_max), c_float(cfg.y_min), c_float(cfg.y_max), c_float(cfg.z_min),
        c_float(cfg.z_max), c_float(cfg.voxel_x), c_float(cfg.voxel_y), c_float(cfg.voxel_z),
        c_int(cfg.grid_x), c_int(cfg.grid_y),
        _p(out["point_features"]), _p(out["point_index_in_voxel"]), _p(out["coords"]),
        _p(out["point_num_in_voxel"]), ctypes.byref(pn), ctypes.byref(ptn))
    assert rc == 0
    out["pillar_num"], out["point_num"] = pn.value, ptn.value
    return out


def window_partition(coords, voxel_num, cfg, which):
    coords = _i32(coords)
    mp, mw, mv = cfg.max_pillars_num, cfg.max_win_num, cfg.max_voxel_num_per_win
    wx, wy, wz = cfg.win_shapes[which]
    sx, sy, sz = cfg.shifts[which]
    out = dict(global_index=np.zeros((mw, mv), np.int32), coors_in_win=np.zeros((mw, mv, 3), np.int32),
               voxel_num_in_win=np.zeros((mw,), np.int32), coors_in_win_2d=np.zeros((mp, 3), np.int32),
               coors_in_win_x_y=np.zeros((mp, 2), np.float32))
    wn = c_int(0)
    rc = lib().oracle_window_partition(
        _p(coords), c_int(int(voxel_num)), c_int(mp), c_int(mw), c_int(mv),
        c_int(cfg.grid_x), c_int(cfg.grid_y), c_int(cfg.grid_z), c_int(wx), c_int(wy), c_int(wz),
        c_int(sx), c_int(sy), c_int(sz),
        _p(out["global_index"]), _p(out["coors_in_win"]), _p(out["voxel_num_in_win"]), ctypes.byref(wn),
        _p(out["coors_in_win_2d"]), _p(out["coors_in_win_x_y"]))
    assert rc == 0
    out["win_num"] = wn.value
    return out


def get_set(global_index, coors_in_win, voxel_num_in_win, win_num, cfg, which):
    mw, mv, S, H = cfg.max_win_num, cfg.max_voxel_num_per_win, cfg.voxel_num_set, cfg.num_heads
    wx, wy, wz = cfg.win_shapes[which]
    out = dict(global_index_in_set=np.zeros((2, mw, S), np.int32), set_voxel_mask=np.zeros((2, mw, S), np.float32),
               mask_expand_0=np.zeros((mw, H, S), np.float32), mask_expand_1=np.zeros((mw, H, S), np.float32))
    sn = c_int(0)
    rc = lib().oracle_get_set(
        _p(_i32(global_index)), _p(_i32(coors_in_win)), _p(_i32(voxel_num_in_win)), c_int(int(win_num)),
        c_int(S), c_int(mw), c_int(mv), c_int(wx), c_int(wy), c_int(wz), c_int(H),
        _p(out["global_index_in_set"]), _p(out["set_voxel_mask"]), ctypes.byref(sn),
        _p(out["mask_expand_0"]), _p(out["mask_expand_1"]))
    assert rc == 0
    out["set_num"] = sn.value
    return out


def gelu(x, voxel_num):
    x = _f32(x)
    out = np.zeros_like(x)
    lib().oracle_gelu(_p(x), c_int(int(voxel_num)), c_int(x.shape[0]), c_int(x.shape[1]), _p(out))
    return out


def layer_norm(x, voxel_num, gamma, beta, eps=0.0, residual=None):
    x = _f32(x)
    out = np.zeros_like(x)
    res = _f32(residual) if residual is not None else None
    lib().oracle_layer_norm(_p(x), _p(res) if res is not None else None, c_int(int(voxel_num)), c_int(x.shape[0]),
                            c_int(x.shape[1]), _p(_f32(gamma)), _p(_f32(beta)), c_float(eps), _p(out))
    return out


def filter_box(scores, classes, xs, ys, center, center_z, angle, dim, cfg):
    K = cfg.max_top_k
    boxes = np.zeros((K, 9), np.float32)
    kept = np.zeros((K,), np.int32)
    valid = c_int(0)
    lib().oracle_filter_box(_p(_f32(scores)), _p(_i32(classes)), _p(_i32(xs)), _p(_i32(ys)), _p(_f32(center)),
                            _p(_f32(center_z)), _p(_f32(angle)), _p(_f32(dim)), c_int(K),
                            c_float(cfg.x_min), c_float(cfg.x_max), c_float(cfg.y_min), c_float(cfg.y_max),
                            c_float(cfg.z_min), c_float(cfg.z_max), c_float(cfg.voxel_x), c_float(cfg.voxel_y),
                            c_float(cfg.score_threshold), _p(boxes), ctypes.byref(valid), _p(kept))
    return boxes, valid.value, kept[: valid.value].copy()


def get_value_by_index(x, pos, idx, set_num, axis):
    x, pos, idx = _f32(x), _f32(pos), _i32(idx)
    _, max_sets, S = idx.shape
    C = x.shape[1]
    q = np.zeros((max_sets, S, C), np.float32)
    k = np.zeros_like(q)
    v = np.zeros_like(q)
    lib().oracle_get_value_by_index(_p(x), _p(pos), _p(idx), c_int(int(set_num)), c_int(max_sets), c_int(S),
                                    c_int(C), c_int(axis), _p(q), _p(k), _p(v))
    return q, k, v


def map_set_feature2voxel(feat, idx, set_num, axis, max_pillars):
    feat, idx = _f32(feat), _i32(idx)
    _, max_sets, S = idx.shape
    C = feat.shape[-1]
    out = np.zeros((max_pillars, C), np.float32)
    lib().oracle_map_set_feature2voxel(_p(feat), _p(idx), c_int(int(set_num)), c_int(max_sets), c_int(S), c_int(C),
                                       c_int(axis), c_int(max_pillars), _p(out))
    return out


def set_attention(q, k, v, mask, n_sets, w_in, b_in, w_out, b_out, heads=8, threads=None):
    """`threads`: the sets are independent, so the serial C routine is called on slices of them from a thread pool
    (ctypes releases the GIL) -- only to keep the large parity cases short; the arithmetic per set is unchanged."""
    q, k, v, mask = _f32(q), _f32(k), _f32(v), _f32(mask)
    _, S, C = q.shape
    out = np.zeros_like(q)
    w_in, b_in, w_out, b_out = _f32(w_in), _f32(b_in), _f32(w_out), _f32(b_out)
    fn = lib().oracle_set_attention
    n_sets = int(n_sets)

    def run(a, b):
        rc = fn(_p(q[a:b]), _p(k[a:b]), _p(v[a:b]), _p(mask[a:b]), c_int(b - a), c_int(S), c_int(C),
                c_int(heads), _p(w_in), _p(b_in), _p(w_out), _p(b_out), _p(out[a:b]))
        assert rc == 0
    if threads is None:
        threads = min(os.cpu_count() or 1, 16) if n_sets >= 256 else 1
    if threads <= 1 or n_sets < 2 * threads:
        run(0, n_sets)
    else:
        from concurrent.futures import ThreadPoolExecutor
        cuts = [n_sets * i // threads for i in range(threads + 1)]
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda ab: run(*ab), zip(cuts[:-1], cuts[1:])))
    return out


def torch_scatter_max(point_features, point_index_in_voxel, point_num_in_voxel, voxel_num):
    f, piv, pnv = _f32(point_features), _i32(point_index_in_voxel), _i32(point_num_in_voxel)
    max_points, F = f.shape
    max_pillars, npv = piv.shape
    mp = np.zeros((max_points, F), np.float32)
    mv = np.zeros((max_pillars, F), np.float32)
    lib().oracle_torch_scatter_max(_p(f), _p(piv), _p(pnv), c_int(int(voxel_num)), c_int(max_points), c_int(max_pillars),
                                   c_int(npv), c_int(F), _p(mp), _p(mv))
    return mp, mv


def map2bev(voxel_features, coords, voxel_num, gx, gy):
    f, co = _f32(voxel_features), _i32(coords)
    max_pillars, C = f.shape
    out = np.zeros((gy, gx, C), np.float32)
    lib().oracle_map2bev(_p(f), _p(co), c_int(int(voxel_num)), c_int(max_pillars), c_int(C), c_int(gx), c_int(gy), _p(out))
    return out


def center_head_topk(heatmap, center, center_z, dim, rot, K=500):
    """The post-process graph of src/dsvt-ai-trt.cpp:1471-1691, literally: sigmoid, TopK(K) per class over H*W, TopK(K) over
    the nc*K survivors, index arithmetic, gathers, exp / atan.  Both TopKs are stable (score descending, earlier element
    first on ties: for the flattened second stage that is class-major, i.e. ascending flat index).  float32 throughout.
    heatmap [nc,H,W] logits, center [2,H,W], center_z [1,H,W], dim [3,H,W], rot [2,H,W] (cos, sin)."""
    hm = _f32(heatmap)
    nc, H, W = hm.shape
    HW = H * W
    s = (np.float32(1.0) / (np.float32(1.0) + np.exp(-hm, dtype=np.float32))).reshape(nc, HW)       # :1479
    # stage 1 orders by the LOGIT so that sigmoid ties (saturation) do not reorder what the GPU selects on; where the
    # sigmoids are distinct this is the same order
    inds = np.stack([np.argsort(-hm.reshape(nc, HW)[c], kind="stable")[:K] for c in range(nc)])      # :1513  [nc,K]
    sc1 = np.take_along_axis(s, inds, axis=1)
    lg1 = np.take_along_axis(hm.reshape(nc, HW), inds, axis=1)
    ind2 = np.argsort(-lg1.reshape(-1), kind="stable")[:K]                                            # :1563
    scores = sc1.reshape(-1)[ind2]
    classes = (ind2 // K).astype(np.int32)                                                            # :1570
    cell = inds.reshape(-1)[ind2]                                                                     # :1592
    ys, xs = (cell // W).astype(np.int32), (cell % W).astype(np.int32)                                # :1543-1547
    c = _f32(center).reshape(2, HW)
    d = np.exp(_f32(dim).reshape(3, HW), dtype=np.float32)                                            # :1489
    r = _f32(rot).reshape(2, HW)
    return dict(scores=scores.astype(np.float32), classes=classes, xs=xs, ys=ys,
                center=np.stack([c[0, cell], c[1, cell]], axis=1), center_z=_f32(center_z).reshape(HW)[cell],
                angle=np.arctan(r[1, cell] / r[0, cell]).astype(np.float32), dim=np.stack([d[0, cell], d[1, cell], d[2, cell]], axis=1))


def nms(boxes, n, nms_thresh):
    """nms_cpu (include/helper.h:257-283) restated in C: input rows of the surviving boxes, in output order."""
    boxes = _f32(boxes)
    keep = np.zeros(max(int(n), 1), np.int32)
    k = lib().oracle_nms(_p(boxes), c_int(int(n)), c_float(nms_thresh), _p(keep))
    return keep[:k].copy()


def nms_iou(boxes, i, j):
    fn = lib().oracle_nms_iou
    fn.restype = ctypes.c_float
    return float(fn(_p(_f32(boxes)), c_int(int(i)), c_int(int(j))))
