"""GPU parity: the sm_100a kernels, called through the C ABI (include/dsvt_b200.h), against the CPU oracle on
identical inputs.  Integer / index outputs must be BIT-EXACT; float outputs within the stated tolerance."""
import importlib

import numpy as np
import pytest
import torch

from conftest import pad_points
from oracle import cpu

pytestmark = pytest.mark.gpu

capi = None


@pytest.fixture(scope="module", autouse=True)
def _capi():
    global capi
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    yield


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def voxelise(points, n, cfg, poison=True):
    vox = capi.Points2Features(cfg)
    if poison:   # outputs must be fully defined by the kernel (zero tails included)
        for t in (vox.point_features, vox.point_index_in_voxel, vox.coords, vox.point_num_in_voxel):
            t.fill_(-7 if t.dtype == torch.int32 else float("nan"))
    vox(dev(pad_points(points, cfg.max_points_num))[None], torch.tensor([n], dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    return vox


def check_voxeliser(points, n, cfg, feat_tol=1e-5):
    vox = voxelise(points, n, cfg)
    o = cpu.points2features(pad_points(points, cfg.max_points_num), n, cfg)
    assert int(vox.pillar_num[0]) == o["pillar_num"]
    assert int(vox.point_num[0]) == o["point_num"]
    assert np.array_equal(vox.coords[0].cpu().numpy(), o["coords"])
    assert np.array_equal(vox.point_num_in_voxel[0].cpu().numpy(), o["point_num_in_voxel"])
    assert np.array_equal(vox.point_index_in_voxel[0].cpu().numpy(), o["point_index_in_voxel"])
    f = vox.point_features[0].cpu().numpy()
    assert not np.isnan(f).any()
    # channels 0..3 are copies of the kept input points -> bit exact; 4..9 are f32/f64 arithmetic
    assert np.array_equal(f[:, :4], o["point_features"][:, :4])
    assert np.abs(f - o["point_features"]).max() <= feat_tol
    return vox, o


# ------------------------------------------------------------------------------------------------
def test_voxeliser_reference_frame(frame0, cfgs, kat):
    vox, o = check_voxeliser(frame0, len(frame0), cfgs.REFERENCE)
    assert int(vox.pillar_num[0]) == kat["000000"]["pillars"] == 5504
    assert int(vox.point_num[0]) == kat["000000"]["kept_points"]


@pytest.mark.parametrize("gen,n", [("ring_lidar", 200000), ("uniform_disc", 200000), ("ring_lidar", 50000)])
def test_voxeliser_waymo_shape(pkg, cfgs, gen, n):
    cfg = cfgs.WAYMO.with_(max_pillars_num=110000)
    pts = getattr(pkg.synth, gen)(n, seed=1)
    check_voxeliser(pts, n, cfg)


def test_voxeliser_pillar_030(pkg, cfgs):
    check_voxeliser(pkg.synth.ring_lidar(200000, seed=2), 200000, cfgs.WAYMO_030)


def test_voxeliser_edge_cases(cfgs):
    cfg = cfgs.REFERENCE
    rng = np.random.default_rng(5)
    # empty cloud
    vox, _ = check_voxeliser(np.zeros((0, 4), np.float32), 0, cfg)
    assert int(vox.pillar_num[0]) == 0 and float(vox.point_features.abs().sum()) == 0
    # every point out of range (z too high / on the exclusive upper edges)
    pts = np.array([[0, 0, 3.0, 1], [74.88, 0, 0, 1], [0, 74.88, 0, 1], [-74.89, 0, 0, 1], [0, 0, -5.01, 1]], np.float32)
    vox, _ = check_voxeliser(pts, len(pts), cfg)
    assert int(vox.pillar_num[0]) == 0
    # inclusive lower edges, duplicates, one pillar with far more than 48 points (keeps the lowest indices)
    heavy = np.tile(np.array([[10.01, -3.33, 0.5, 9]], np.float32), (5000, 1))
    heavy[:, 3] = np.arange(5000)
    heavy[:, 2] = rng.uniform(-4, 2, 5000)
    pts = np.concatenate([np.array([[-74.88, -74.88, -5.0, 3]], np.float32), heavy,
                          rng.uniform(-70, 70, (3000, 4)).astype(np.float32) * np.array([1, 1, 0.03, 1], np.float32)])
    check_voxeliser(pts, len(pts), cfg)
    # points_size smaller than the data present and larger than the capacity
    check_voxeliser(pts, 100, cfg)
    big = rng.uniform(-70, 70, (60000, 4)).astype(np.float32) * np.array([1, 1, 0.03, 1], np.float32)
    check_voxeliser(big, 60000, cfg)   # > MAX_POINTS_NUM: truncated like the reference (points2Features.cu:716)
    # pillar and row capacities exceeded -> deterministic clamp, no out-of-bounds writes
    vox, o = check_voxeliser(big, 50000, cfg.with_(max_pillars_num=1000))
    assert int(vox.pillar_num[0]) == 1000 and 1000 <= int(vox.point_num[0]) < 2000   # dropped pillars emit no rows
    vox, o = check_voxeliser(big, 50000, cfg.with_(max_pillars_num=40000, max_points_num_voxel_filter=1500))
    assert int(vox.point_num[0]) == 1500 and int(vox.pillar_num[0]) > 1500          # later pillars keep 0 points


def test_voxeliser_batched(pkg, cfgs):
    """Batch extension: B frames in one launch == B single-frame launches."""
    cfg = cfgs.WAYMO
    B = 3
    clouds = [pkg.synth.ring_lidar(180000 - 7000 * i, seed=10 + i) for i in range(B)]
    pts = np.stack([pad_points(c, cfg.max_points_num) for c in clouds])
    sizes = torch.tensor([len(c) for c in clouds], dtype=torch.int32, device="cuda")
    vb = capi.Points2Features(cfg, batch=B)
    vb(dev(pts), sizes)
    torch.cuda.synchronize()
    for i, c in enumerate(clouds):
        o = cpu.points2features(pad_points(c, cfg.max_points_num), len(c), cfg)
        assert int(vb.pillar_num[i]) == o["pillar_num"] and int(vb.point_num[i]) == o["point_num"]
        assert np.array_equal(vb.coords[i].cpu().numpy(), o["coords"])
        assert np.array_equal(vb.point_index_in_voxel[i].cpu().numpy(), o["point_index_in_voxel"])
        assert np.abs(vb.point_features[i].cpu().numpy() - o["point_features"]).max() <= 1e-5


def test_voxeliser_deterministic(pkg, cfgs):
    cfg = cfgs.WAYMO
    pts = pkg.synth.uniform_disc(250000, seed=4)
    a = voxelise(pts, len(pts), cfg.with_(max_pillars_num=110000))
    ref = [t.clone() for t in (a.point_features, a.point_index_in_voxel, a.coords, a.point_num_in_voxel)]
    for _ in range(3):
        b = voxelise(pts, len(pts), cfg.with_(max_pillars_num=110000))
        for r, t in zip(ref, (b.point_features, b.point_index_in_voxel, b.coords, b.point_num_in_voxel)):
            assert torch.equal(r, t)


# ------------------------------------------------------------------------------------------------
def run_partition(o, cfg, which):
    wp = capi.WindowPartition(cfg, which)
    for t in (wp.global_index, wp.coors_in_win, wp.voxel_num_in_win, wp.coors_in_win_2d):
        t.fill_(-7)
    wp.coors_in_win_x_y.fill_(float("nan"))
    wp(dev(o["coords"])[None], torch.tensor([o["pillar_num"]], dtype=torch.int32, device="cuda"))
    gs = capi.GetSet(cfg, which)
    gs.global_index_in_set.fill_(-7)
    for t in (gs.set_voxel_mask, gs.mask_expand_0, gs.mask_expand_1):
        t.fill_(float("nan"))
    gs(wp.global_index, wp.coors_in_win, wp.voxel_num_in_win, wp.win_num)
    torch.cuda.synchronize()
    return wp, gs


def check_partition(o, cfg, which):
    wp, gs = run_partition(o, cfg, which)
    owp = cpu.window_partition(o["coords"], o["pillar_num"], cfg, which)
    assert int(wp.win_num[0]) == owp["win_num"]
    for name in ("global_index", "coors_in_win", "voxel_num_in_win", "coors_in_win_2d", "coors_in_win_x_y"):
        assert np.array_equal(getattr(wp, name)[0].cpu().numpy(), owp[name]), name
    ogs = cpu.get_set(owp["global_index"], owp["coors_in_win"], owp["voxel_num_in_win"], owp["win_num"], cfg, which)
    assert int(gs.set_num[0]) == ogs["set_num"]
    assert np.array_equal(gs.global_index_in_set[0].cpu().numpy(), ogs["global_index_in_set"])
    for name in ("set_voxel_mask", "mask_expand_0", "mask_expand_1"):   # compare bit patterns (0 / -FLT_MAX)
        assert np.array_equal(getattr(gs, name)[0].cpu().numpy().view(np.uint32), ogs[name].view(np.uint32)), name
    return owp, ogs


def test_partition_reference_frame(frame0, cfgs, kat):
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    for which, tag in ((0, "win12"), (1, "win24_shift6")):
        owp, ogs = check_partition(o, cfg, which)
        assert owp["win_num"] == kat["000000"][tag]["windows"] and ogs["set_num"] == kat["000000"][tag]["sets"]


def test_partition_window_capacity_overflow(frame0, cfgs):
    """max_voxel_num_per_win below the cells of a window: overfull windows keep their LOWEST voxel ids (the outcome of the
    reference's loop executed serially, windowPartition.cu:303) -- deterministic and bit-exact against the oracle."""
    cfg = cfgs.REFERENCE.with_(max_voxel_num_per_win=20)
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    for which in (0, 1):
        owp, _ = check_partition(o, cfg, which)
        assert (owp["voxel_num_in_win"] == 20).sum() > 10          # the overflow path really ran


@pytest.mark.parametrize("S", [24, 36, 48])
def test_partition_waymo_shape(pkg, cfgs, S):
    cfg = cfgs.WAYMO.with_(voxel_num_set=S, max_pillars_num=110000, max_win_num=8192)
    pts = pkg.synth.uniform_disc(200000, seed=3)
    o = cpu.points2features(pad_points(pts, cfg.max_points_num), len(pts), cfg)
    for which in (0, 1):
        check_partition(o, cfg, which)


def test_get_set_edge_cases(cfgs):
    cfg = cfgs.REFERENCE.with_(max_win_num=64)
    rng = np.random.default_rng(2)

    def run(gi, cw, vn, wn, which=0):
        gs = capi.GetSet(cfg, which)
        gs.global_index_in_set.fill_(-7)
        gs(dev(gi)[None], dev(cw)[None], dev(vn)[None], torch.tensor([wn], dtype=torch.int32, device="cuda"))
        torch.cuda.synchronize()
        ref = cpu.get_set(gi, cw, vn, wn, cfg, which)
        assert int(gs.set_num[0]) == ref["set_num"]
        assert np.array_equal(gs.global_index_in_set[0].cpu().numpy(), ref["global_index_in_set"])
        assert np.array_equal(gs.mask_expand_1[0].cpu().numpy().view(np.uint32), ref["mask_expand_1"].view(np.uint32))
        return ref

    mw, mv = cfg.max_win_num, cfg.max_voxel_num_per_win
    gi = np.zeros((mw, mv), np.int32)
    cw = np.zeros((mw, mv, 3), np.int32)
    vn = np.zeros((mw,), np.int32)
    assert run(gi, cw, vn, 0)["set_num"] == 0                               # no windows at all
    # windows with 1, exactly S, S+1 and the maximum 144 voxels (12x12 window completely full)
    nid = 0
    for w, n in enumerate([1, 36, 37, 144, 35, 72, 73]):
        cells = rng.permutation(144)[:n]                                   # unsorted on purpose
        cw[w, :n, 1], cw[w, :n, 2] = cells // 12, cells % 12
        gi[w, :n] = nid + rng.permutation(n)
        vn[w] = n
        nid += n
    ref = run(gi, cw, vn, 7)
    assert ref["set_num"] == 1 + 1 + 2 + 4 + 1 + 2 + 3
    # more sets than max_win_num: clamped, nothing written out of bounds
    for w in range(mw):
        n = 40
        cells = rng.permutation(144)[:n]
        cw[w, :n, 1], cw[w, :n, 2] = cells // 12, cells % 12
        gi[w, :n] = w * 40 + np.arange(n)
        vn[w] = n
    assert run(gi, cw, vn, mw)["set_num"] == mw


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("V,cap", [(5504, 10000), (0, 10000), (10000, 10000), (30611, 40000), (1, 7)])
def test_gelu(V, cap):
    rng = np.random.default_rng(V)
    x = (rng.standard_normal((cap, 384)) * 4).astype(np.float32)
    x[0, :8] = [0.0, -0.0, 1e-30, -1e-30, 30.0, -30.0, 88.0, -88.0]
    out = torch.full((cap, 384), float("nan"), device="cuda")
    capi.gelu(dev(x), torch.tensor([V], dtype=torch.int32, device="cuda"), out=out)
    ref = cpu.gelu(x, V)
    got = out.cpu().numpy()
    assert np.all(got[V:] == 0)
    # tolerance: f32 logistic form vs the reference's double tanh -- 1e-6 abs + 2e-6 rel (<< 1e-3 bar)
    assert np.all(np.abs(got - ref) <= 1e-6 + 2e-6 * np.abs(ref))


@pytest.mark.parametrize("V,cap,C", [(5504, 10000, 192), (0, 10000, 192), (10000, 10000, 192), (30611, 40000, 192),
                                     (77, 100, 96)])
@pytest.mark.parametrize("with_res", [False, True])
def test_layer_norm(V, cap, C, with_res):
    rng = np.random.default_rng(V + C)
    x = (rng.standard_normal((cap, C)) * 3 + 0.7).astype(np.float32)
    res = rng.standard_normal((cap, C)).astype(np.float32) if with_res else None
    gamma, beta = rng.standard_normal(C).astype(np.float32), rng.standard_normal(C).astype(np.float32)
    out = torch.full((cap, C), float("nan"), device="cuda")
    capi.layer_norm(dev(x), torch.tensor([V], dtype=torch.int32, device="cuda"), dev(gamma), dev(beta), eps=0.0,
                    residual=dev(res) if with_res else None, out=out)
    ref = cpu.layer_norm(x, V, gamma, beta, 0.0, residual=res)
    got = out.cpu().numpy()
    assert np.all(got[V:] == 0)
    # tolerance: the reference sums sequentially in f32, we tree-reduce; 2e-5 abs on O(1..10) outputs
    assert np.abs(got - ref).max() <= 2e-5 if V else True


@pytest.mark.parametrize("n_stages", [1, 2, 3])
def test_layer_norm_chain(n_stages):
    """Chained LayerNorms == the same LayerNorm plugins applied one after the other (oracle composition)."""
    rng = np.random.default_rng(n_stages)
    V, cap, C = 16566, 20000, 192
    x = (rng.standard_normal((cap, C)) * 2).astype(np.float32)
    res = [(rng.standard_normal((cap, C))).astype(np.float32) if i != 1 or n_stages == 3 else None for i in range(n_stages)]
    gam = [rng.standard_normal(C).astype(np.float32) for _ in range(n_stages)]
    bet = [rng.standard_normal(C).astype(np.float32) for _ in range(n_stages)]
    ref = x
    for i in range(n_stages):
        ref = cpu.layer_norm(ref, V, gam[i], bet[i], 0.0, residual=res[i])
    vn = torch.tensor([V], dtype=torch.int32, device="cuda")
    out = torch.full((cap, C), float("nan"), device="cuda")
    capi.layer_norm_chain(dev(x), vn, [(dev(r) if r is not None else None, dev(g), dev(b)) for r, g, b in zip(res, gam, bet)],
                          0.0, out=out)
    got = out.cpu().numpy()
    assert np.all(got[V:] == 0)
    assert np.abs(got - ref).max() <= 5e-5 * n_stages
    # and bit-identical to the un-fused kernels of this library
    seq = dev(x)
    for i in range(n_stages):
        seq = capi.layer_norm(seq, vn, dev(gam[i]), dev(bet[i]), 0.0, residual=dev(res[i]) if res[i] is not None else None)
    assert torch.equal(seq, out)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_filter_box(pkg, cfgs, seed):
    cfg = cfgs.REFERENCE
    arrs = list(pkg.synth.head_candidates(cfg.max_top_k, seed=seed))
    if seed == 1:
        arrs[0][:] = 0.0            # nothing passes the score threshold
    if seed == 2:
        arrs[0][:] = 0.9            # everything in range passes
        arrs[4][:] = 0.5
        arrs[5][:] = 0.0
    boxes = torch.full((1, cfg.max_top_k, 9), float("nan"), device="cuda")
    _, valid = capi.filter_box(cfg, *[dev(a)[None] for a in arrs], boxes=boxes)
    ref_boxes, ref_valid, _ = cpu.filter_box(*arrs, cfg)
    assert int(valid[0]) == ref_valid
    assert np.array_equal(boxes[0].cpu().numpy(), ref_boxes)    # one FMA + copies: bit exact


# ------------------------------------------------------------------------------------------------
def _attn_inputs(n_sets, max_sets, S=36, C=192, H=8, seed=0):
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((max_sets, S, C)).astype(np.float32)
    k = q.copy()
    v = rng.standard_normal((max_sets, S, C)).astype(np.float32)
    mask = np.zeros((max_sets, H, S), np.float32)
    for s in range(max_sets):
        pad = rng.permutation(S - 1)[: rng.integers(0, S - 1)] + 1
        mask[s, :, pad] = -np.finfo(np.float32).max
    w = dict(w_in=(rng.standard_normal((3 * C, C)) * 0.06).astype(np.float32),
             b_in=(rng.standard_normal(3 * C) * 0.1).astype(np.float32),
             w_out=(rng.standard_normal((C, C)) * 0.06).astype(np.float32),
             b_out=(rng.standard_normal(C) * 0.1).astype(np.float32))
    return q, k, v, mask, w


# DSVT_ATTN_FP32: same f32 arithmetic, different summation order.  DSVT_ATTN_FP32_TC (3): FP16 hi+lo split operands on
# tcgen05 (2^-22 relative per product), FP32 accumulate -- held to the SAME tolerance as the CUDA-core FP32 path.
ATTN_TOL = {0: 2e-5, 2: 1e-2, 3: 2e-5, 4: 1e-2}


@pytest.mark.parametrize("S", [24, 36, 48])
def test_set_attention_plugin_form(S):
    n_sets, max_sets = 37, 50
    q, k, v, mask, w = _attn_inputs(n_sets, max_sets, S=S, seed=S)
    k = (k + 0.25 * np.random.default_rng(1).standard_normal(k.shape)).astype(np.float32)   # q != k in general
    W = capi.AttentionWeights(w["w_in"], w["b_in"], w["w_out"], w["b_out"])
    out = torch.full((max_sets, S, 192), float("nan"), device="cuda")
    capi.set_attention(W, dev(q), dev(k), dev(v), dev(mask), torch.tensor([n_sets], dtype=torch.int32, device="cuda"),
                       out=out, precision=0)
    ref = cpu.set_attention(q, k, v, mask, n_sets, **w)
    got = out.cpu().numpy()
    assert np.all(got[n_sets:] == 0)
    assert np.abs(got[:n_sets] - ref[:n_sets]).max() <= ATTN_TOL[0]
    # set_num == NULL -> all max_sets sets are computed, as the reference graph does
    out2 = capi.set_attention(W, dev(q), dev(k), dev(v), dev(mask), None, precision=0)
    ref2 = cpu.set_attention(q, k, v, mask, max_sets, **w)
    assert np.abs(out2.cpu().numpy() - ref2).max() <= ATTN_TOL[0]


def test_set_attention_golden(attention_case):
    c = attention_case
    W = capi.AttentionWeights(c["w_in"], c["b_in"], c["w_out"], c["b_out"])
    out = capi.set_attention(W, dev(c["q"]), dev(c["k"]), dev(c["v"]), dev(c["mask"]), None, precision=0)
    assert np.abs(out.cpu().numpy() - c["out"]).max() <= 2e-5      # torch MHA fixture


@pytest.mark.parametrize("precision", [0, 2, 3, 4])
def test_set_attention_fused_frame(frame0, cfgs, precision):
    """Fused gather + attention + scatter on the reference frame == oracle gather -> attention -> scatter
    (precision 2 = the USE_FP16 configuration's single fused tcgen05 kernel, BASELINE.json configs[2], tolerance 1e-2)."""
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V = o["pillar_num"]
    rng = np.random.default_rng(11)
    x = np.zeros((cfg.max_pillars_num, 192), np.float32)
    pos = np.zeros_like(x)
    x[:V] = rng.standard_normal((V, 192))
    pos[:V] = rng.standard_normal((V, 192)) * 0.5
    _, _, _, _, w = _attn_inputs(1, 1, seed=5)
    W = capi.AttentionWeights(w["w_in"], w["b_in"], w["w_out"], w["b_out"])
    for which in (0, 1):
        owp = cpu.window_partition(o["coords"], V, cfg, which)
        ogs = cpu.get_set(owp["global_index"], owp["coors_in_win"], owp["voxel_num_in_win"], owp["win_num"], cfg, which)
        ns = ogs["set_num"]
        for axis in (0, 1):
            q, k, v = cpu.get_value_by_index(x, pos, ogs["global_index_in_set"], ns, axis)
            a = cpu.set_attention(q, k, v, ogs["mask_expand_0"], ns, **w)
            ref = cpu.map_set_feature2voxel(a, ogs["global_index_in_set"], ns, axis, cfg.max_pillars_num)
            out = torch.full((cfg.max_pillars_num, 192), float("nan"), device="cuda")
            capi.set_attention_fused(W, dev(x), dev(pos), dev(ogs["global_index_in_set"]), dev(ogs["mask_expand_0"]),
                                     torch.tensor([ns], dtype=torch.int32, device="cuda"),
                                     torch.tensor([V], dtype=torch.int32, device="cuda"), axis, out=out,
                                     precision=precision)
            got = out.cpu().numpy()
            assert np.all(got[V:] == 0)
            err = np.abs(got - ref).max()
            assert err <= ATTN_TOL[precision], err
            if precision:
                continue
            # standalone gather / scatter plugins
            gq, gk, gv = capi.get_value_by_index(dev(x), dev(pos), dev(ogs["global_index_in_set"]),
                                                 torch.tensor([ns], dtype=torch.int32, device="cuda"), axis)
            assert np.array_equal(gq.cpu().numpy(), q) and np.array_equal(gk.cpu().numpy(), k)
            assert np.array_equal(gv.cpu().numpy(), v)
            sc = capi.map_set_feature2voxel(dev(a), dev(ogs["global_index_in_set"]),
                                            torch.tensor([ns], dtype=torch.int32, device="cuda"), axis,
                                            cfg.max_pillars_num)
            assert np.array_equal(sc.cpu().numpy(), ref)


@pytest.mark.parametrize("precision", [0, 2, 3, 4])
def test_set_attention_fused_set_capacity_overflow(frame0, cfgs, precision):
    """max_win_num below the frame's set count: GetSet drops the sets beyond the capacity, their voxels belong to no set and
    the reference's scatter into a zero-filled tensor (mapSetFeature2voxel.cu:312) leaves those rows exactly 0 -- also
    when the workspace still holds rows of an earlier call."""
    cfg = cfgs.REFERENCE.with_(max_win_num=300)            # frame 0 has 454 sets at 12x12
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V = o["pillar_num"]
    rng = np.random.default_rng(12)
    x = np.zeros((cfg.max_pillars_num, 192), np.float32)
    pos = np.zeros_like(x)
    x[:V] = rng.standard_normal((V, 192))
    pos[:V] = rng.standard_normal((V, 192)) * 0.5
    _, _, _, _, w = _attn_inputs(1, 1, seed=6)
    W = capi.AttentionWeights(w["w_in"], w["b_in"], w["w_out"], w["b_out"])
    owp = cpu.window_partition(o["coords"], V, cfg, 0)
    ogs = cpu.get_set(owp["global_index"], owp["coors_in_win"], owp["voxel_num_in_win"], owp["win_num"], cfg, 0)
    ns = ogs["set_num"]
    assert ns == 300
    ws_bytes = capi.set_attention_workspace_bytes(1, cfg.max_win_num, 36, 192, 8, cfg.max_pillars_num, precision)
    ws = torch.full((max(ws_bytes, 256),), 0x7F, dtype=torch.uint8, device="cuda")      # stale workspace: large finite floats
    for axis in (0, 1):
        q, k, v = cpu.get_value_by_index(x, pos, ogs["global_index_in_set"], ns, axis)
        a = cpu.set_attention(q, k, v, ogs["mask_expand_0"], ns, **w)
        ref = cpu.map_set_feature2voxel(a, ogs["global_index_in_set"], ns, axis, cfg.max_pillars_num)
        covered = np.zeros(cfg.max_pillars_num, bool)
        covered[np.unique(ogs["global_index_in_set"][axis, :ns])] = True
        assert (~covered[:V]).sum() > 500
        out = torch.full((cfg.max_pillars_num, 192), float("nan"), device="cuda")
        capi.set_attention_fused(W, dev(x), dev(pos), dev(ogs["global_index_in_set"]), dev(ogs["mask_expand_0"]),
                                 torch.tensor([ns], dtype=torch.int32, device="cuda"),
                                 torch.tensor([V], dtype=torch.int32, device="cuda"), axis, out=out, precision=precision,
                                 workspace=ws if ws_bytes else None)
        got = out.cpu().numpy()
        assert np.all(got[~covered] == 0), "voxels in no set must be exactly zero"
        assert np.abs(got - ref).max() <= ATTN_TOL[precision]


@pytest.mark.parametrize("precision", [3, 4])
def test_set_attention_plan_reuse(frame0, cfgs, precision):
    """A plan built once per (partition, axis) and passed to several layers gives exactly the stateless result."""
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V = o["pillar_num"]
    rng = np.random.default_rng(3)
    owp = cpu.window_partition(o["coords"], V, cfg, 1)
    ogs = cpu.get_set(owp["global_index"], owp["coors_in_win"], owp["voxel_num_in_win"], owp["win_num"], cfg, 1)
    ns_t = torch.tensor([ogs["set_num"]], dtype=torch.int32, device="cuda")
    v_t = torch.tensor([V], dtype=torch.int32, device="cuda")
    idx, mask = dev(ogs["global_index_in_set"]), dev(ogs["mask_expand_0"])
    for axis in (0, 1):
        plan = capi.set_attention_plan(idx, mask, ns_t, axis, cfg.max_pillars_num)
        for layer in range(2):                       # two layers, different weights / activations, same plan
            x = np.zeros((cfg.max_pillars_num, 192), np.float32); pos = np.zeros_like(x)
            x[:V] = rng.standard_normal((V, 192)); pos[:V] = rng.standard_normal((V, 192)) * 0.5
            _, _, _, _, w = _attn_inputs(1, 1, seed=20 + layer)
            W = capi.AttentionWeights(w["w_in"], w["b_in"], w["w_out"], w["b_out"])
            a = capi.set_attention_fused(W, dev(x), dev(pos), idx, mask, ns_t, v_t, axis, precision=precision)
            b = capi.set_attention_fused(W, dev(x), dev(pos), idx, mask, ns_t, v_t, axis, precision=precision, plan=plan)
            assert torch.equal(a, b)
            ref = capi.set_attention_fused(W, dev(x), dev(pos), idx, mask, ns_t, v_t, axis, precision=0)
            assert (a[:V] - ref[:V]).abs().max().item() <= ATTN_TOL[precision]


# ------------------------------------------------------------------------------------------------
# FP16 tensor-core configuration (reference: USE_FP16, params.h:332).  Tolerance 1e-2 abs (BASELINE.json
# configs[2]); FP16 operands carry 11-bit significands, the measured error is ~1e-3.
@pytest.mark.parametrize("precision", [2, 3, 4])
@pytest.mark.parametrize("n_sets", [1, 2, 3, 4, 100, 454, 1450])
def test_set_attention_fused_fp16_tensor_cores(n_sets, precision):
    """Tensor-core paths (2: fused FP16 kernel, 3: FP32-accurate split GEMM pipeline, 4: FP16 GEMM pipeline) against the
    CUDA-core FP32 kernel on synthetic set partitions, and against the CPU oracle on the small cases."""
    tol = ATTN_TOL.get(precision, 1e-2)
    rng = np.random.default_rng(n_sets)
    max_sets, S, C, H = max(8, n_sets + 3), 36, 192, 8
    sizes = rng.integers(1, S + 1, n_sets)        # every voxel belongs to exactly one set (as getSet guarantees)
    V = int(sizes.sum())
    max_pillars = V + 37
    x = np.zeros((max_pillars, C), np.float32)
    pos = np.zeros_like(x)
    x[:V] = rng.standard_normal((V, C))
    pos[:V] = rng.standard_normal((V, C)) * 0.5
    idx = np.zeros((2, max_sets, S), np.int32)
    mask = np.zeros((max_sets, H, S), np.float32)
    perm = rng.permutation(V)
    start = 0
    for s in range(n_sets):                       # ranks with repeats, like DSVT eq.(3)
        n = int(sizes[s])
        members = np.sort(perm[start:start + n])
        start += n
        r = (np.arange(S) * n) // S
        idx[0, s] = members[r]
        idx[1, s] = members[::-1][r]
        mask[s, :, 1:][:, r[1:] == r[:-1]] = -np.finfo(np.float32).max
    _, _, _, _, w = _attn_inputs(1, 1, seed=3)
    W = capi.AttentionWeights(w["w_in"], w["b_in"], w["w_out"], w["b_out"])
    ns_t = torch.tensor([n_sets], dtype=torch.int32, device="cuda")
    v_t = torch.tensor([V], dtype=torch.int32, device="cuda")
    for axis in (0, 1):
        ref = capi.set_attention_fused(W, dev(x), dev(pos), dev(idx), dev(mask), ns_t, v_t, axis, precision=0)
        out = torch.full((max_pillars, C), float("nan"), device="cuda")
        capi.set_attention_fused(W, dev(x), dev(pos), dev(idx), dev(mask), ns_t, v_t, axis, out=out,
                                 precision=precision)
        torch.cuda.synchronize()
        got, want = out.cpu().numpy(), ref.cpu().numpy()
        touched = np.unique(idx[axis, :n_sets])
        assert not np.isnan(got[touched]).any()
        assert np.all(got[V:] == 0)
        err = np.abs(got[touched] - want[touched]).max()
        assert err <= (2 * tol if precision == 3 else tol), err     # both sides carry their own error vs the oracle
        # and against the CPU oracle directly (every size: the oracle's sets run on a thread pool)
        q, k, v = cpu.get_value_by_index(x, pos, idx, n_sets, axis)
        a = cpu.set_attention(q, k, v, mask, n_sets, **w)
        o = cpu.map_set_feature2voxel(a, idx, n_sets, axis, max_pillars)
        assert np.abs(got[touched] - o[touched]).max() <= tol


def _synthetic_partition(rng, n_sets, max_sets, S, H=8):
    """Sets of 1..S distinct voxels with DSVT eq.(3)-style repeats; every voxel belongs to exactly one set."""
    sizes = rng.integers(1, S + 1, n_sets)
    V = int(sizes.sum())
    idx = np.zeros((2, max_sets, S), np.int32)
    mask = np.zeros((max_sets, H, S), np.float32)
    perm, start = rng.permutation(V), 0
    for s in range(n_sets):
        n = int(sizes[s])
        members = np.sort(perm[start:start + n]); start += n
        r = (np.arange(S) * n) // S
        idx[0, s] = members[r]; idx[1, s] = members[::-1][r]
        mask[s, :, 1:][:, r[1:] == r[:-1]] = -np.finfo(np.float32).max
    return idx, mask, V


@pytest.mark.parametrize("S", [24, 48])
def test_set_attention_pipeline_other_set_sizes(S):
    """GEMM pipeline (precision 3) for set sizes 24 / 48 (BASELINE.json config 5) against the CPU oracle chain."""
    rng = np.random.default_rng(S)
    n_sets, max_sets, C = 30, 40, 192
    idx, mask, V = _synthetic_partition(rng, n_sets, max_sets, S)
    max_pillars = V + 19
    x = np.zeros((max_pillars, C), np.float32); pos = np.zeros_like(x)
    x[:V] = rng.standard_normal((V, C)); pos[:V] = rng.standard_normal((V, C)) * 0.5
    _, _, _, _, w = _attn_inputs(1, 1, seed=9)
    W = capi.AttentionWeights(w["w_in"], w["b_in"], w["w_out"], w["b_out"])
    ns_t = torch.tensor([n_sets], dtype=torch.int32, device="cuda"); v_t = torch.tensor([V], dtype=torch.int32, device="cuda")
    for axis in (0, 1):
        out = torch.full((max_pillars, C), float("nan"), device="cuda")
        capi.set_attention_fused(W, dev(x), dev(pos), dev(idx), dev(mask), ns_t, v_t, axis, out=out, precision=3)
        q, k, v = cpu.get_value_by_index(x, pos, idx, n_sets, axis)
        a = cpu.set_attention(q, k, v, mask, n_sets, **w)
        ref = cpu.map_set_feature2voxel(a, idx, n_sets, axis, max_pillars)
        got = out.cpu().numpy()
        assert np.all(got[V:] == 0)
        assert np.abs(got[:V] - ref[:V]).max() <= ATTN_TOL[3]


@pytest.mark.parametrize("precision", [3, 2])
def test_set_attention_fused_batched(precision):
    """batch = 3 frames with different set / voxel counts in one launch sequence == the three single-frame results."""
    rng = np.random.default_rng(77)
    B, max_sets, S, C = 3, 64, 36, 192
    parts = [_synthetic_partition(rng, n, max_sets, S) for n in (50, 7, 33)]
    max_pillars = max(p[2] for p in parts) + 40
    x = np.zeros((B, max_pillars, C), np.float32); pos = np.zeros_like(x)
    for b, (_, _, V) in enumerate(parts):
        x[b, :V] = rng.standard_normal((V, C)); pos[b, :V] = rng.standard_normal((V, C)) * 0.5
    idx = np.stack([p[0] for p in parts]); mask = np.stack([p[1] for p in parts])
    ns = torch.tensor([50, 7, 33], dtype=torch.int32, device="cuda")
    vn = torch.tensor([p[2] for p in parts], dtype=torch.int32, device="cuda")
    _, _, _, _, w = _attn_inputs(1, 1, seed=4)
    W = capi.AttentionWeights(w["w_in"], w["b_in"], w["w_out"], w["b_out"])
    out = torch.full((B, max_pillars, C), float("nan"), device="cuda")
    capi.set_attention_fused(W, dev(x), dev(pos), dev(idx), dev(mask), ns, vn, 1, out=out, precision=precision)
    for b, (_, _, V) in enumerate(parts):
        one = capi.set_attention_fused(W, dev(x[b]), dev(pos[b]), dev(idx[b]), dev(mask[b]), ns[b:b + 1], vn[b:b + 1], 1,
                                       precision=precision)
        assert torch.equal(out[b, :V], one[:V]) and float(out[b, V:].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------
# next #3: TorchScatterMaxPlugin / Map2BevPlugin (pure max / copy arithmetic: bit exact)
@pytest.mark.parametrize("F", [96, 192])
def test_torch_scatter_max(frame0, cfgs, F):
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V, Pc = o["pillar_num"], o["point_num"]
    rng = np.random.default_rng(F)
    feat = np.zeros((cfg.max_points_num_voxel_filter, F), np.float32)
    feat[:Pc] = rng.standard_normal((Pc, F)) * 3
    feat[:Pc][rng.random((Pc, F)) < 0.01] = -2000000.0           # below the reference's -1000000 start value
    ref_mp, ref_mv = cpu.torch_scatter_max(feat, o["point_index_in_voxel"], o["point_num_in_voxel"], V)
    piv, pnv = dev(o["point_index_in_voxel"]), dev(o["point_num_in_voxel"].reshape(-1))
    v_t = torch.tensor([V], dtype=torch.int32, device="cuda")
    for point_num in (None, torch.tensor([Pc], dtype=torch.int32, device="cuda")):     # reference form / with the row count
        mp = torch.full((cfg.max_points_num_voxel_filter, F), float("nan"), device="cuda")
        mv = torch.full((cfg.max_pillars_num, F), float("nan"), device="cuda")
        capi.torch_scatter_max(dev(feat), piv, pnv, v_t, point_num, max_point=mp, max_voxel=mv)
        assert np.array_equal(mp.cpu().numpy(), ref_mp) and np.array_equal(mv.cpu().numpy(), ref_mv)
    # no pillars at all: both outputs are all zero
    mp, mv = capi.torch_scatter_max(dev(feat), piv, pnv, torch.zeros(1, dtype=torch.int32, device="cuda"),
                                    torch.zeros(1, dtype=torch.int32, device="cuda"))
    assert float(mp.abs().max()) == 0.0 and float(mv.abs().max()) == 0.0


def test_torch_scatter_max_batched_waymo(pkg, cfgs):
    cfg = cfgs.WAYMO.with_(max_points_num=120000, max_points_num_voxel_filter=120000)
    B, F = 2, 96
    rng = np.random.default_rng(0)
    feats = np.zeros((B, cfg.max_points_num_voxel_filter, F), np.float32)
    pivs, pnvs, Vs, Pcs, refs = [], [], [], [], []
    for b in range(B):
        pts = pkg.synth.ring_lidar(100000, seed=10 + b)
        o = cpu.points2features(pad_points(pts, cfg.max_points_num), len(pts), cfg)
        feats[b, : o["point_num"]] = rng.standard_normal((o["point_num"], F))
        pivs.append(o["point_index_in_voxel"]); pnvs.append(o["point_num_in_voxel"].reshape(-1))
        Vs.append(o["pillar_num"]); Pcs.append(o["point_num"])
        refs.append(cpu.torch_scatter_max(feats[b], o["point_index_in_voxel"], o["point_num_in_voxel"], o["pillar_num"]))
    mp, mv = capi.torch_scatter_max(dev(feats), dev(np.stack(pivs)), dev(np.stack(pnvs)),
                                    torch.tensor(Vs, dtype=torch.int32, device="cuda"),
                                    torch.tensor(Pcs, dtype=torch.int32, device="cuda"))
    for b in range(B):
        assert np.array_equal(mp[b].cpu().numpy(), refs[b][0]) and np.array_equal(mv[b].cpu().numpy(), refs[b][1])


def test_map2bev(frame0, cfgs):
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V = o["pillar_num"]
    rng = np.random.default_rng(2)
    x = np.zeros((cfg.max_pillars_num, 192), np.float32)
    x[:V] = rng.standard_normal((V, 192))
    ref = cpu.map2bev(x, o["coords"], V, cfg.grid_x, cfg.grid_y)
    out = torch.full((cfg.grid_y, cfg.grid_x, 192), float("nan"), device="cuda")
    capi.map2bev(dev(x), dev(o["coords"]), torch.tensor([V], dtype=torch.int32, device="cuda"), cfg.grid_x, cfg.grid_y, out=out)
    got = out.cpu().numpy()
    assert np.array_equal(got, ref)
    assert int((np.abs(got).sum(axis=2) > 0).sum()) == V             # one occupied cell per pillar
    # a coordinate outside the grid is ignored (the reference would write out of bounds)
    bad = o["coords"].copy(); bad[0, 2] = cfg.grid_y + 5
    out2 = capi.map2bev(dev(x), dev(bad), torch.tensor([V], dtype=torch.int32, device="cuda"), cfg.grid_x, cfg.grid_y)
    assert np.array_equal(out2.cpu().numpy(), cpu.map2bev(x, bad, V, cfg.grid_x, cfg.grid_y))
