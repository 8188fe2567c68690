"""The TensorRT plugin shells, driven through the C harness the way TensorRT drives them
(creator lookup -> createPlugin(fields) -> getOutputDimensions -> getWorkspaceSize -> enqueue ->
serialize -> deserializePlugin -> clone), against the CPU oracle."""
import importlib
import struct

import numpy as np
import pytest
import torch

from conftest import pad_points
from oracle import cpu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def plg():
    return importlib.import_module("dsvt-ai-trt_b200.plugins")


@pytest.fixture(scope="module")
def lib(plg):
    return plg.PluginLibrary()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def i32(v):
    return torch.tensor([v], dtype=torch.int32, device="cuda")


def roundtrip(lib, plugin):
    """serialize -> deserialize -> clone: returns the re-created plugin."""
    blob = plugin.serialize()
    again = lib.deserialize(plugin.type, blob)
    assert again.serialize() == blob
    c = again.clone()
    assert c.serialize() == blob and c.type == plugin.type and c.version == "1"
    return c, blob


def test_voxel_generator_plugin(lib, plg, frame0, cfgs):
    cfg = cfgs.REFERENCE
    p = plg.add_voxel_generator(lib, cfg.max_points_num, cfg.max_points_num_voxel_filter, cfg.max_pillars_num, 4, 10,
                                cfg.max_num_points_per_voxel, cfg.x_min, cfg.x_max, cfg.y_min, cfg.y_max, cfg.z_min,
                                cfg.z_max, cfg.voxel_x, cfg.voxel_y, cfg.voxel_z, cfg.grid_x, cfg.grid_y, cfg.grid_z)
    assert p.type == "Points2FeaturesPlugin" and p.nb_outputs == 6
    p2, blob = roundtrip(lib, p)
    # serialised layout of the reference: 6 x i32, 9 x f32 (xmin,xmax,ymin,ymax,zmin,zmax,vx,vy,vz), 3 x i32
    assert blob == struct.pack("<6i9f3i", 50000, 30000, 10000, 4, 10, 48, -74.88, 74.88, -74.88, 74.88, -5.0, 3.0,
                               0.32, 0.32, 8.0, 468, 468, 1)
    pts = pad_points(frame0, cfg.max_points_num)
    o = cpu.points2features(pts, len(frame0), cfg)
    for plugin in (p, p2):
        outs = plugin.enqueue([dev(pts)[None], i32(len(frame0))], poison=-3)
        torch.cuda.synchronize()
        assert [tuple(t.shape) for t in outs] == [(1, 30000, 10), (1, 10000, 48), (1, 10000, 4), (1, 10000, 1), (1,), (1,)]
        assert [t.dtype for t in outs] == [torch.float32] + [torch.int32] * 5
        assert int(outs[4][0]) == 5504 and int(outs[5][0]) == o["point_num"]
        assert np.array_equal(outs[2][0].cpu().numpy(), o["coords"])
        assert np.array_equal(outs[1][0].cpu().numpy(), o["point_index_in_voxel"])
        assert np.array_equal(outs[3][0, :, 0].cpu().numpy(), o["point_num_in_voxel"])
        assert np.abs(outs[0][0].cpu().numpy() - o["point_features"]).max() <= 1e-5


def test_window_partition_and_get_set_plugins(lib, plg, frame0, cfgs):
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    for which in (0, 1):
        wp = plg.add_window_partition(lib, cfg.max_win_num, cfg.max_voxel_num_per_win,
                                      (cfg.grid_x, cfg.grid_y, cfg.grid_z), cfg.win_shapes[which], cfg.shifts[which])
        gs = plg.add_get_set_op(lib, cfg.max_win_num, cfg.max_voxel_num_per_win, cfg.voxel_num_set, cfg.win_shapes[which])
        wp, wblob = roundtrip(lib, wp)
        gs, gblob = roundtrip(lib, gs)
        w = cfg.win_shapes[which]
        s = cfg.shifts[which]
        assert wblob == struct.pack("<11i", 468, 468, 1, *w, *s, 800, 576)
        assert gblob == struct.pack("<6i", 36, 800, 576, *w)
        wouts = wp.enqueue([dev(o["coords"])[None], i32(o["pillar_num"])], poison=-3)
        gouts = gs.enqueue(wouts[:4], poison=-3)
        torch.cuda.synchronize()
        assert [tuple(t.shape) for t in wouts] == [(1, 800, 576), (1, 800, 576, 3), (1, 800), (1,), (1, 10000, 3), (1, 10000, 2)]
        assert [tuple(t.shape) for t in gouts] == [(1, 2, 800, 36), (1, 2, 800, 36), (1,), (1, 800, 8, 36), (1, 800, 8, 36)]
        assert [t.dtype for t in gouts] == [torch.int32, torch.float32, torch.int32, torch.float32, torch.float32]
        owp = cpu.window_partition(o["coords"], o["pillar_num"], cfg, which)
        ogs = cpu.get_set(owp["global_index"], owp["coors_in_win"], owp["voxel_num_in_win"], owp["win_num"], cfg, which)
        assert np.array_equal(wouts[0][0].cpu().numpy(), owp["global_index"])
        assert np.array_equal(wouts[1][0].cpu().numpy(), owp["coors_in_win"])
        assert np.array_equal(wouts[5][0].cpu().numpy(), owp["coors_in_win_x_y"])
        assert int(gouts[2][0]) == ogs["set_num"]
        assert np.array_equal(gouts[0][0].cpu().numpy(), ogs["global_index_in_set"])
        assert np.array_equal(gouts[1][0].cpu().numpy().view(np.uint32), ogs["set_voxel_mask"].view(np.uint32))
        assert np.array_equal(gouts[3][0].cpu().numpy().view(np.uint32), ogs["mask_expand_0"].view(np.uint32))
        assert np.array_equal(gouts[4][0].cpu().numpy().view(np.uint32), ogs["mask_expand_1"].view(np.uint32))


def test_gelu_and_layer_norm_plugins(lib, plg, cfgs):
    cfg = cfgs.REFERENCE
    rng = np.random.default_rng(0)
    V = 5504
    g = plg.add_gelu_op(lib, cfg.max_pillars_num, 384)
    g, blob = roundtrip(lib, g)
    assert blob == struct.pack("<2i", 10000, 384)
    x = (rng.standard_normal((1, cfg.max_pillars_num, 384)) * 3).astype(np.float32)
    (out,) = g.enqueue([dev(x), i32(V)], poison=float("nan"))
    ref = cpu.gelu(x[0], V)
    assert np.all(np.abs(out[0].cpu().numpy() - ref) <= 1e-6 + 2e-6 * np.abs(ref))

    gamma, beta = rng.standard_normal(192).astype(np.float32), rng.standard_normal(192).astype(np.float32)
    ln = plg.add_layer_norm_op(lib, cfg.max_pillars_num, 192, gamma, beta, eps=1e-5)
    ln2, blob = roundtrip(lib, ln)
    # eps is never forwarded by the reference helper ("pes" quirk): serialised eps must be 0.0
    assert blob[:16] == struct.pack("<3if", 10000, 192, 192, 0.0)
    assert blob[16:] == gamma.tobytes() + beta.tobytes()
    x = (rng.standard_normal((1, cfg.max_pillars_num, 192)) * 2 + 1).astype(np.float32)
    ref = cpu.layer_norm(x[0], V, gamma, beta, 0.0)
    for plugin in (ln, ln2):
        (out,) = plugin.enqueue([dev(x), i32(V)], poison=float("nan"))
        assert np.abs(out[0].cpu().numpy() - ref).max() <= 2e-5
    assert "pes" in lib.field_names("LayerNormPlugin") and "eps" not in lib.field_names("LayerNormPlugin")


def test_filter_box_plugin(lib, plg, pkg, cfgs):
    cfg = cfgs.REFERENCE
    fb = plg.add_filter_box_by_score_op(lib, cfg.max_top_k, cfg.x_min, cfg.x_max, cfg.y_min, cfg.y_max, cfg.z_min,
                                        cfg.z_max, cfg.voxel_x, cfg.voxel_y, cfg.voxel_z, cfg.score_threshold)
    fb, blob = roundtrip(lib, fb)
    assert blob == struct.pack("<i10f", 500, -74.88, 74.88, -74.88, 74.88, -5.0, 3.0, 0.32, 0.32, 8.0, 0.3)
    sc, cl, xs, ys, ce, cz, an, dm = pkg.synth.head_candidates(cfg.max_top_k, seed=4)
    ins = [dev(sc)[None], dev(cl)[None], dev(xs)[None], dev(ys)[None], dev(ce)[None, None], dev(cz)[None, None, :, None],
           dev(an)[None, None, :, None], dev(dm)[None, None]]
    boxes, valid = fb.enqueue(ins, poison=float("nan"))
    rb, rv, _ = cpu.filter_box(sc, cl, xs, ys, ce, cz, an, dm, cfg)
    assert tuple(boxes.shape) == (1, 500, 9) and int(valid[0]) == rv
    assert np.array_equal(boxes[0].cpu().numpy(), rb)


def test_set_attention_plugin(lib, plg, attention_case):
    c = attention_case
    n, S, C = c["q"].shape
    p = plg.add_set_attention_op(lib, n, S, C, 8, c["w_in"], c["b_in"], c["w_out"], c["b_out"], precision=0)
    p2, blob = roundtrip(lib, p)
    assert len(blob) == 6 * 4 + (4 * C * C + 4 * C) * 4       # 5 x i32 config + the connected input count (4 / 5) + weights
    assert struct.unpack("<6i", blob[:24]) == (n, S, C, 8, 0, 4)
    for plugin in (p, p2):
        (out,) = plugin.enqueue([dev(c["q"])[None], dev(c["k"])[None], dev(c["v"])[None], dev(c["mask"])[None]],
                                poison=float("nan"))
        assert np.abs(out[0].cpu().numpy() - c["out"]).max() <= 2e-5
    # optional 5th input: valid set count
    (out,) = p.enqueue([dev(c["q"])[None], dev(c["k"])[None], dev(c["v"])[None], dev(c["mask"])[None], i32(2)],
                       poison=float("nan"))
    got = out[0].cpu().numpy()
    assert np.abs(got[:2] - c["out"][:2]).max() <= 2e-5 and np.all(got[2:] == 0)
    # the optional input survives clone() and serialize -> deserialize (TensorRT clones after configurePlugin and
    # deserialises at run time): the clone must not fall back to "all max_set_num sets"
    for again in (p.clone(), lib.deserialize(p.type, p.serialize())):
        assert struct.unpack("<6i", again.serialize()[:24])[5] == 5


@pytest.mark.parametrize("precision", [0, 2, 3, 4])
def test_set_attention_fused_plugin(lib, plg, precision):
    """SetAttentionFusedPlugin (gather + MHA + scatter in one node) through the plugin interface: creator fields,
    serialise / deserialise round trip, getWorkspaceSize, enqueue -- against the CPU oracle chain."""
    rng = np.random.default_rng(5)
    n_sets, max_sets, S, C, H = 40, 48, 36, 192, 8
    sizes = rng.integers(1, S + 1, n_sets)
    V = int(sizes.sum())
    max_pillars = V + 21
    x = np.zeros((max_pillars, C), np.float32); pos = np.zeros_like(x)
    x[:V] = rng.standard_normal((V, C)); pos[:V] = rng.standard_normal((V, C)) * 0.5
    idx = np.zeros((2, max_sets, S), np.int32)
    mask = np.zeros((max_sets, H, S), np.float32)
    perm, start = rng.permutation(V), 0
    for s in range(n_sets):
        n = int(sizes[s])
        members = np.sort(perm[start:start + n]); start += n
        r = (np.arange(S) * n) // S
        idx[0, s] = members[r]; idx[1, s] = members[::-1][r]
        mask[s, :, 1:][:, r[1:] == r[:-1]] = -np.finfo(np.float32).max
    w_in = (rng.standard_normal((3 * C, C)) * 0.06).astype(np.float32); b_in = (rng.standard_normal(3 * C) * 0.1).astype(np.float32)
    w_out = (rng.standard_normal((C, C)) * 0.06).astype(np.float32); b_out = (rng.standard_normal(C) * 0.1).astype(np.float32)
    tol = {0: 2e-5, 3: 2e-5}.get(precision, 1e-2)
    for axis in (0, 1):
        p = plg.add_set_attention_fused_op(lib, max_sets, S, C, H, max_pillars, axis, w_in, b_in, w_out, b_out,
                                           precision=precision)
        p2, blob = roundtrip(lib, p)
        assert blob[:28] == struct.pack("<7i", max_sets, S, C, H, precision, max_pillars, axis)
        assert len(blob) == 7 * 4 + (4 * C * C + 4 * C) * 4 + 8            # + (has_norm = 0, eps) trailer
        q, k, v = cpu.get_value_by_index(x, pos, idx, n_sets, axis)
        a = cpu.set_attention(q, k, v, mask, n_sets, w_in, b_in, w_out, b_out)
        ref = cpu.map_set_feature2voxel(a, idx, n_sets, axis, max_pillars)
        for plugin in (p, p2):
            (out,) = plugin.enqueue([dev(x)[None], dev(pos)[None], dev(idx)[None], dev(mask)[None], i32(n_sets), i32(V)],
                                    poison=float("nan"))
            got = out[0].cpu().numpy()
            assert np.all(got[V:] == 0)
            assert np.abs(got[:V] - ref[:V]).max() <= tol
        if precision in (3, 4):     # planned form: SetAttentionPlanPlugin output as the optional 7th input
            pl = plg.add_set_attention_plan_op(lib, max_sets, S, H, max_pillars, axis)
            pl, pblob = roundtrip(lib, pl)
            assert pblob == struct.pack("<5i", max_sets, S, H, max_pillars, axis)
            (plan,) = pl.enqueue([dev(idx)[None], dev(mask)[None], i32(n_sets)])
            (out7,) = p.enqueue([dev(x)[None], dev(pos)[None], dev(idx)[None], dev(mask)[None], i32(n_sets), i32(V), plan],
                                poison=float("nan"))
            assert torch.equal(out7, out)


@pytest.mark.parametrize("precision", [3, 4])
def test_set_attention_plugin_tensor_cores(lib, plg, attention_case, precision):
    """The true drop-in for multHeadAttention() (q, k, v pre-gathered by GetValueByIndexPlugin) on the tensor-core pipeline:
    SetAttentionPlugin with precision DSVT_ATTN_FP32_TC / DSVT_ATTN_FP16_GEMM against the torch-MHA golden case and, with
    the optional set count, against the oracle on a padded frame-shaped input."""
    c = attention_case
    n, S, C = c["q"].shape
    tol = 2e-5 if precision == 3 else 1e-2
    p = plg.add_set_attention_op(lib, n, S, C, 8, c["w_in"], c["b_in"], c["w_out"], c["b_out"], precision=precision)
    p2, _ = roundtrip(lib, p)
    for plugin in (p, p2):
        (out,) = plugin.enqueue([dev(c["q"])[None], dev(c["k"])[None], dev(c["v"])[None], dev(c["mask"])[None]],
                                poison=float("nan"))
        assert np.abs(out[0].cpu().numpy() - c["out"]).max() <= tol
    rng = np.random.default_rng(8)
    n_sets, max_sets = 150, 200
    q = rng.standard_normal((max_sets, S, C)).astype(np.float32)
    k = (q + 0.25 * rng.standard_normal(q.shape)).astype(np.float32)
    v = rng.standard_normal((max_sets, S, C)).astype(np.float32)
    mask = np.zeros((max_sets, 8, S), np.float32)
    for s_ in range(max_sets):
        mask[s_, :, rng.permutation(S - 1)[: rng.integers(0, S - 1)] + 1] = -np.finfo(np.float32).max
    p = plg.add_set_attention_op(lib, max_sets, S, C, 8, c["w_in"], c["b_in"], c["w_out"], c["b_out"], precision=precision)
    (out,) = p.enqueue([dev(q)[None], dev(k)[None], dev(v)[None], dev(mask)[None], i32(n_sets)], poison=float("nan"))
    ref = cpu.set_attention(q, k, v, mask, n_sets, c["w_in"], c["b_in"], c["w_out"], c["b_out"])
    got = out[0].cpu().numpy()
    assert np.all(got[n_sets:] == 0)
    assert np.abs(got[:n_sets] - ref[:n_sets]).max() <= tol


def test_set_attention_fused_plugin_with_norm(lib, plg, frame0, cfgs):
    """SetAttentionFusedPlugin with the optional norm fields = GetValueByIndex -> multHeadAttention -> MapSetFeature2Voxel ->
    kSUM -> LayerNormPlugin (src/dsvt-ai-trt.cpp:653-676) as ONE node, on the reference frame, against the oracle chain."""
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V, C = o["pillar_num"], 192
    owp = cpu.window_partition(o["coords"], V, cfg, 0)
    ogs = cpu.get_set(owp["global_index"], owp["coors_in_win"], owp["voxel_num_in_win"], owp["win_num"], cfg, 0)
    ns, idx, mask = ogs["set_num"], ogs["global_index_in_set"], ogs["mask_expand_0"]
    rng = np.random.default_rng(21)
    x = np.zeros((cfg.max_pillars_num, C), np.float32); pos = np.zeros_like(x)
    x[:V] = rng.standard_normal((V, C)); pos[:V] = rng.standard_normal((V, C)) * 0.5
    w_in = (rng.standard_normal((3 * C, C)) * 0.06).astype(np.float32); b_in = (rng.standard_normal(3 * C) * 0.1).astype(np.float32)
    w_out = (rng.standard_normal((C, C)) * 0.06).astype(np.float32); b_out = (rng.standard_normal(C) * 0.1).astype(np.float32)
    gamma = (1 + 0.1 * rng.standard_normal(C)).astype(np.float32); beta = (0.1 * rng.standard_normal(C)).astype(np.float32)
    for axis in (0, 1):
        p = plg.add_set_attention_fused_op(lib, cfg.max_win_num, 36, C, 8, cfg.max_pillars_num, axis, w_in, b_in, w_out, b_out,
                                           precision=3, norm_weight=gamma, norm_bias=beta)
        p2, blob = roundtrip(lib, p)
        assert len(blob) == 7 * 4 + (4 * C * C + 4 * C) * 4 + 8 + 2 * C * 4
        q, k, v = cpu.get_value_by_index(x, pos, idx, ns, axis)
        a = cpu.set_attention(q, k, v, mask, ns, w_in, b_in, w_out, b_out)
        y = cpu.map_set_feature2voxel(a, idx, ns, axis, cfg.max_pillars_num)
        ref = cpu.layer_norm(y, V, gamma, beta, 0.0, residual=x)
        for plugin in (p, p2):
            (out,) = plugin.enqueue([dev(x)[None], dev(pos)[None], dev(idx)[None], dev(mask)[None], i32(ns), i32(V)],
                                    poison=float("nan"))
            got = out[0].cpu().numpy()
            assert np.all(got[V:] == 0)
            assert np.abs(got - ref).max() <= 1e-4


@pytest.mark.parametrize("n_stages", [1, 2, 3])
def test_layer_norm_chain_plugin(lib, plg, n_stages):
    """LayerNormChainPlugin == n_stages x (addElementWise(kSUM) + LayerNormPlugin), bit-identical to the single plugins."""
    rng = np.random.default_rng(n_stages)
    mp, C, V = 700, 192, 555
    x = np.zeros((mp, C), np.float32); x[:V] = rng.standard_normal((V, C))
    res = [rng.standard_normal((mp, C)).astype(np.float32) for _ in range(n_stages)]
    gam = (1 + 0.1 * rng.standard_normal((n_stages, C))).astype(np.float32)
    bet = (0.1 * rng.standard_normal((n_stages, C))).astype(np.float32)
    p = plg.add_layer_norm_chain_op(lib, mp, C, gam, bet)
    p, blob = roundtrip(lib, p)
    assert blob[:16] == struct.pack("<3if", mp, C, n_stages, 0.0)
    (out,) = p.enqueue([dev(x)[None], i32(V)] + [dev(r)[None] for r in res], poison=float("nan"))
    ref, single = x, dev(x)[None]
    for s_ in range(n_stages):
        ref = cpu.layer_norm(ref, V, gam[s_], bet[s_], 0.0, residual=res[s_])
        ln = plg.add_layer_norm_op(lib, mp, C, gam[s_], bet[s_])
        (single,) = ln.enqueue([single + dev(res[s_])[None], i32(V)])
    got = out[0].cpu().numpy()
    assert np.all(got[V:] == 0) and np.abs(got - ref).max() <= 2e-5
    assert torch.equal(out, single)


@pytest.mark.parametrize("K,N,act,n_ln", [(192, 384, 1, 0), (384, 192, 0, 0), (384, 192, 0, 2), (384, 192, 0, 3), (192, 192, 2, 0)])
def test_linear_plugin(lib, plg, K, N, act, n_ln):
    """LinearPlugin: a TensorRT FullyConnected layer of the 3-D backbone (+ GELU / ReLU, or + the LayerNorm chain behind the
    FFN) as one node, against float64."""
    rng = np.random.default_rng(K + N + act + n_ln)
    mr, rows = 1300, 1000
    x = np.zeros((mr, K), np.float32); x[:rows] = rng.standard_normal((rows, K))
    W = (rng.standard_normal((N, K)) * 0.06).astype(np.float32); b = (rng.standard_normal(N) * 0.05).astype(np.float32)
    res = [rng.standard_normal((mr, N)).astype(np.float32) for _ in range(n_ln)]
    gam = (1 + 0.1 * rng.standard_normal((n_ln, N))).astype(np.float32)
    bet = (0.1 * rng.standard_normal((n_ln, N))).astype(np.float32)
    p = plg.add_linear_op(lib, mr, K, N, W, b, activation=act, ln_weights=gam if n_ln else None, ln_bias=bet if n_ln else None)
    p, blob = roundtrip(lib, p)
    assert blob[:24] == struct.pack("<6i", mr, K, N, act, 3, n_ln)
    (out,) = p.enqueue([dev(x)[None], i32(rows)] + [dev(r)[None] for r in res], poison=float("nan"))
    y = x[:rows].astype(np.float64) @ W.T.astype(np.float64) + b
    if act == 1:
        y = (0.5 + 0.5 * np.tanh(y * (0.035677408136300125 * y * y + 0.7978845608028654))) * y
    elif act == 2:
        y = np.maximum(y, 0.0)
    for s_ in range(n_ln):
        y = y + res[s_][:rows]
        mu = y.mean(1, keepdims=True)
        y = (y - mu) / np.sqrt(((y - mu) ** 2).mean(1, keepdims=True)) * gam[s_] + bet[s_]
    got = out[0].cpu().numpy()
    assert np.all(got[rows:] == 0)
    assert np.abs(got[:rows] - y).max() <= 5e-5


@pytest.mark.parametrize("n_ln", [2, 3])
def test_ffn_fused_plugin(lib, plg, n_ln):
    """FfnFusedPlugin: FC 192->384, GELU, FC 384->192 and the LayerNorm chain behind it as one node, against float64 and its own
    serialised clone."""
    rng = np.random.default_rng(40 + n_ln)
    mr, rows, C, Fh = 1300, 1000, 192, 384
    x = np.zeros((mr, C), np.float32); x[:rows] = rng.standard_normal((rows, C))
    W1 = (rng.standard_normal((Fh, C)) * 0.07).astype(np.float32); b1 = (rng.standard_normal(Fh) * 0.05).astype(np.float32)
    W2 = (rng.standard_normal((C, Fh)) * 0.05).astype(np.float32); b2 = (rng.standard_normal(C) * 0.05).astype(np.float32)
    res = [rng.standard_normal((mr, C)).astype(np.float32) for _ in range(n_ln)]
    gam = (1 + 0.1 * rng.standard_normal((n_ln, C))).astype(np.float32)
    bet = (0.1 * rng.standard_normal((n_ln, C))).astype(np.float32)
    p = plg.add_ffn_fused_op(lib, mr, W1, b1, W2, b2, gam, bet)
    p, blob = roundtrip(lib, p)
    assert blob[:8] == struct.pack("<2i", mr, n_ln)
    (out,) = p.enqueue([dev(x)[None], i32(rows)] + [dev(r)[None] for r in res], poison=float("nan"))
    h = x[:rows].astype(np.float64) @ W1.T.astype(np.float64) + b1
    h = (0.5 + 0.5 * np.tanh(h * (0.035677408136300125 * h * h + 0.7978845608028654))) * h
    y = h @ W2.T.astype(np.float64) + b2
    for s_ in range(n_ln):
        y = y + res[s_][:rows]
        mu = y.mean(1, keepdims=True)
        y = (y - mu) / np.sqrt(((y - mu) ** 2).mean(1, keepdims=True)) * gam[s_] + bet[s_]
    got = out[0].cpu().numpy()
    assert np.all(got[rows:] == 0)
    assert np.abs(got[:rows] - y).max() <= 5e-5


def test_vfe_fused_plugin(lib, plg, frame0):
    """VfeFusedPlugin on Points2FeaturesPlugin's own outputs (the reference's sample frame) against the four nodes it replaces,
    run through the C ABI: bit for bit."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    config = importlib.import_module("dsvt-ai-trt_b200.config")
    cfg = config.REFERENCE
    rng = np.random.default_rng(11)
    pts = np.zeros((cfg.max_points_num, 4), np.float32)
    n = min(len(frame0), cfg.max_points_num)
    pts[:n] = frame0[:n]
    vox = capi.Points2Features(cfg)
    vox(dev(pts)[None], i32(n))
    W0 = (rng.standard_normal((96, 10)) * 0.3).astype(np.float32)
    sc0, sh0 = (1 + 0.1 * rng.standard_normal(96)).astype(np.float32), (0.1 * rng.standard_normal(96)).astype(np.float32)
    W1 = (rng.standard_normal((192, 192)) * 0.07).astype(np.float32); b1 = (rng.standard_normal(192) * 0.05).astype(np.float32)
    p = plg.add_vfe_fused_op(lib, cfg.max_points_num_voxel_filter, cfg.max_pillars_num, cfg.max_num_points_per_voxel, W0, sc0, sh0, W1, b1)
    p, blob = roundtrip(lib, p)
    assert blob[:12] == struct.pack("<3i", cfg.max_points_num_voxel_filter, cfg.max_pillars_num, cfg.max_num_points_per_voxel)
    (out,) = p.enqueue([vox.point_features, vox.point_index_in_voxel, vox.pillar_num, vox.point_num], poison=float("nan"))
    pfn0 = capi.SmallLinear(W0, sc0, sh0)
    pfn1 = capi.Linear(W1, b1, precision=capi.DSVT_ATTN_FP32_TC)
    h0 = pfn0(vox.point_features[0], vox.point_num, activation=2, zero_tails=0)
    mp0, _ = capi.torch_scatter_max(h0, vox.point_index_in_voxel[0], vox.point_num_in_voxel[0], vox.pillar_num, vox.point_num)
    h1 = pfn1.rows_concat(h0, mp0, vox.point_num, activation=2, zero_tails=0)
    _, mv1 = capi.torch_scatter_max(h1, vox.point_index_in_voxel[0], vox.point_num_in_voxel[0], vox.pillar_num, vox.point_num)
    torch.cuda.synchronize()
    V = int(vox.pillar_num[0])
    assert V > 1000 and torch.equal(out[0][:V], mv1[:V]) and bool((out[0][V:] == 0).all())


def test_registry_and_formats(lib):
    names = set(lib.registered())
    assert {"Points2FeaturesPlugin", "GetSetPlugin", "GeluPlugin", "LayerNormPlugin", "FilterBoxByScorePlugin",
            "WindowPartitionPlugin", "GetValueByIndexPlugin", "MapSetFeature2VoxelPlugin", "SetAttentionPlugin",
            "SetAttentionFusedPlugin", "SetAttentionPlanPlugin", "TorchScatterMaxPlugin", "Map2BevPlugin",
            "LayerNormChainPlugin", "LinearPlugin", "FfnFusedPlugin", "VfeFusedPlugin"} <= names
