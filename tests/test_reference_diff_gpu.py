"""Differential parity against the REFERENCE'S OWN KERNELS.

oracle/_ref/libref_<plugin>.so are the reference's plugin sources, compiled unmodified for sm_100a
against the in-repo TensorRT scaffold header (oracle/build.py) and linked with the same C harness our
plugins use.  Both sides are created from identical PluginFieldCollections and fed identical device
buffers.  The reference's outputs depend on atomicAdd races (pillar / row / set / box order, which 48
points survive in an overfull pillar); they are compared after the canonicalisation of SURVEY.md
Appendix A.  The libraries are prebuilt in the authoring container and travel with the snapshot;
nothing here reads /root/reference at run time.
"""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, pad_points
from oracle import cpu

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def ref_lib(plg, stem):
    path = os.path.join(REF_DIR, f"libref_{stem}.so")
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (oracle/build.py needs /root/reference)")
    return plg.PluginLibrary(path)


@pytest.fixture(scope="module")
def plg():
    return importlib.import_module("dsvt-ai-trt_b200.plugins")


@pytest.fixture(scope="module")
def ours(plg):
    return plg.PluginLibrary()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def i32(v):
    return torch.tensor([v], dtype=torch.int32, device="cuda")


def make_voxeliser(plg, lib, cfg):
    return plg.add_voxel_generator(lib, cfg.max_points_num, cfg.max_points_num_voxel_filter, cfg.max_pillars_num, 4, 10,
                                   cfg.max_num_points_per_voxel, cfg.x_min, cfg.x_max, cfg.y_min, cfg.y_max,
                                   cfg.z_min, cfg.z_max, cfg.voxel_x, cfg.voxel_y, cfg.voxel_z, cfg.grid_x,
                                   cfg.grid_y, cfg.grid_z)


def canonical_voxels(outs, cfg):
    """-> dict cell -> (count, sorted rows of the pillar's 10-channel features), via coords + point_index_in_voxel."""
    feats, piv, coords, pnv, pn, ptn = [t[0].cpu().numpy() for t in outs]
    V, P = int(pn), int(ptn)
    res = {}
    used_rows = []
    for p in range(V):
        n = int(pnv[p, 0]) if pnv.ndim == 2 else int(pnv[p])
        rows = piv[p, :n]
        used_rows.append(rows)
        f = feats[rows]
        order = np.lexsort((f[:, 3], f[:, 2], f[:, 1], f[:, 0]))
        res[int(coords[p, 2]) * cfg.grid_x + int(coords[p, 3])] = (n, f[order])
        assert coords[p, 0] == 0 and coords[p, 1] == 0
    used = np.concatenate(used_rows) if used_rows else np.zeros(0, np.int64)
    assert len(np.unique(used)) == len(used) == P, "row ids form a permutation of 0..point_num-1"
    assert used.max(initial=-1) == P - 1
    return V, P, res


@pytest.mark.parametrize("case", ["frame0", "ring29k_dense"])
def test_points2features_vs_reference_kernels(plg, ours, pkg, cfgs, frame0, case):
    cfg = cfgs.REFERENCE
    # the reference has no capacity guards (SURVEY A-5): stay below 30000 kept rows / 10000 pillars
    # (ring_lidar(28000) has 10582 pillars: the reference then overwrites its neighbouring tensors)
    if case == "frame0":
        pts = frame0
    else:
        pts = pkg.synth.ring_lidar(29000, seed=3)
        pts[:, :2] *= 0.3          # ~3k pillars, a few hundred of them overfull
    n = len(pts)
    buf = dev(pad_points(pts, cfg.max_points_num))[None]
    mine = make_voxeliser(plg, ours, cfg).enqueue([buf, i32(n)], poison=-3)
    ref = make_voxeliser(plg, ref_lib(plg, "points2Features"), cfg).enqueue([buf, i32(n)], poison=-3)
    torch.cuda.synchronize()
    Vm, Pm, cm = canonical_voxels(mine, cfg)
    Vr, Pr, cr = canonical_voxels(ref, cfg)
    assert (Vm, Pm) == (Vr, Pr)
    assert cm.keys() == cr.keys(), "same set of non-empty pillars"
    npv = cfg.max_num_points_per_voxel
    x0, y0, vs = np.float32(cfg.x_min), np.float32(cfg.y_min), np.float32(cfg.voxel_x)
    n_full = 0
    for cell, (nm, fm) in cm.items():
        nr, fr = cr[cell]
        assert nm == nr
        if nm < npv:
            # same points; xyzi copies bit-exact, offsets within 1e-5 (mean summation order differs)
            assert np.array_equal(fm[:, :4], fr[:, :4])
            assert np.abs(fm - fr).max() <= 1e-5
        else:
            # overfull (or exactly full) pillar: WHICH 48 survive is a race in the reference (SURVEY A-2);
            # require: 48 kept, all inside the pillar, pillar-centre channels consistent
            n_full += 1
            for f in (fm, fr):
                assert np.all(np.floor((f[:, 0] - x0) / vs) == cell % cfg.grid_x)
                assert np.all(np.floor((f[:, 1] - y0) / vs) == cell // cfg.grid_x)
    assert n_full >= 1
    # tails are zero on both sides
    assert float(mine[0][0, Pm:].abs().sum()) == 0 and float(ref[0][0, Pr:].abs().sum()) == 0


def test_window_partition_and_get_set_vs_reference_kernels(plg, ours, cfgs, frame0):
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    coords, V = dev(o["coords"])[None], i32(o["pillar_num"])
    for which in (0, 1):
        args = (cfg.max_win_num, cfg.max_voxel_num_per_win, (cfg.grid_x, cfg.grid_y, cfg.grid_z),
                cfg.win_shapes[which], cfg.shifts[which])
        m = plg.add_window_partition(ours, *args).enqueue([coords, V], poison=-3)
        r = plg.add_window_partition(ref_lib(plg, "windowPartition"), *args).enqueue([coords, V], poison=-3)
        torch.cuda.synchronize()
        assert int(m[3][0]) == int(r[3][0])
        W = int(m[3][0])

        # Window slot order and in-window voxel order are atomicAdd races in the reference (windowPartition.cu:302,
        # :309).  Worse, a voxel that reads the not-yet-published window id 0 sleeps and re-reads through a plain
        # (cacheable, non-volatile) load (:323-331) and then files itself under window slot 0, and voxel_num_in_win is
        # published from a count read earlier (:334-339).  So: slot 0 is skipped, and in every other reference window
        # each filled position must hold a voxel of the ONE matching window of ours, with identical coordinates.
        gi_m, cw_m, vn_m = m[0][0].cpu().numpy(), m[1][0].cpu().numpy(), m[2][0].cpu().numpy()
        gi_r, cw_r = r[0][0].cpu().numpy(), r[1][0].cpu().numpy()
        win_of, coord_of = {}, {}
        for w in range(W):
            for p in range(vn_m[w]):
                win_of[int(gi_m[w, p])] = w
                coord_of[int(gi_m[w, p])] = cw_m[w, p]
        assert len(win_of) == o["pillar_num"]
        seen, matched = set(), 0
        for w in range(1, W):
            mw = win_of[int(gi_r[w, 0])]            # position 0 is written by the thread that opened the window
            assert mw not in seen
            seen.add(mw)
            for p in range(int(vn_m[mw])):
                v = int(gi_r[w, p])
                if v == 0 and win_of[0] != mw:
                    continue                          # position never written: its voxel was misfiled under slot 0
                assert win_of[v] == mw and np.array_equal(cw_r[w, p], coord_of[v])
                matched += 1
            assert np.all(gi_r[w, int(vn_m[mw]):] == 0)
        assert matched >= 0.5 * o["pillar_num"]
        print(f"windowPartition[{which}]: {W} windows, {matched}/{o['pillar_num']} voxels filed correctly by the reference")
        # per-voxel outputs are order independent
        assert torch.equal(m[4], r[4]) and torch.equal(m[5], r[5])

        # getSet on IDENTICAL inputs (ours): set order is a race in the reference -> sort sets
        gargs = (cfg.max_win_num, cfg.max_voxel_num_per_win, cfg.voxel_num_set, cfg.win_shapes[which])
        gm = plg.add_get_set_op(ours, *gargs).enqueue(m[:4], poison=-3)
        gr = plg.add_get_set_op(ref_lib(plg, "getSet"), *gargs).enqueue(m[:4], poison=-3)
        torch.cuda.synchronize()
        ns = int(gm[2][0])
        assert ns == int(gr[2][0])

        def sets(outs):
            idx, msk = outs[0][0].cpu().numpy(), outs[1][0].cpu().numpy().view(np.uint32)
            e0, e1 = outs[3][0].cpu().numpy().view(np.uint32), outs[4][0].cpu().numpy().view(np.uint32)
            order = np.lexsort(idx[0, :ns].T[::-1])
            return idx[:, :ns][:, order], msk[:, :ns][:, order], e0[:ns][order], e1[:ns][order]
        for a, b in zip(sets(gm), sets(gr)):
            assert np.array_equal(a, b)
        for k, (t_m, t_r) in enumerate(zip(gm, gr)):
            if k == 2:
                continue       # set_num
            assert float(t_m[0].flatten()[-36:].abs().sum()) == 0 == float(t_r[0].flatten()[-36:].abs().sum())


def test_rowwise_plugins_vs_reference_kernels(plg, ours, pkg, cfgs):
    cfg = cfgs.REFERENCE
    rng = np.random.default_rng(1)
    V = 5504
    x = (rng.standard_normal((1, cfg.max_pillars_num, 384)) * 3).astype(np.float32)
    gm = plg.add_gelu_op(ours, cfg.max_pillars_num, 384).enqueue([dev(x), i32(V)], poison=float("nan"))[0]
    gr = plg.add_gelu_op(ref_lib(plg, "gelu"), cfg.max_pillars_num, 384).enqueue([dev(x), i32(V)])[0]
    d = (gm - gr).abs()
    assert float((d - 2e-6 * gr.abs()).max()) <= 1e-6        # f32 logistic form vs the reference's double tanh

    gamma, beta = rng.standard_normal(192).astype(np.float32), rng.standard_normal(192).astype(np.float32)
    x = (rng.standard_normal((1, cfg.max_pillars_num, 192)) * 2 + 1).astype(np.float32)
    lm = plg.add_layer_norm_op(ours, cfg.max_pillars_num, 192, gamma, beta).enqueue([dev(x), i32(V)], poison=float("nan"))[0]
    lr = plg.add_layer_norm_op(ref_lib(plg, "layerNorm"), cfg.max_pillars_num, 192, gamma, beta).enqueue([dev(x), i32(V)])[0]
    assert float((lm - lr).abs().max()) <= 2e-5

    # filterBox: the reference launches 512 threads over 500 candidates without a guard (SURVEY A-10); give it
    # inputs carved from 512-element buffers whose tail cannot pass the score threshold
    K = cfg.max_top_k
    sc, cl, xs, ys, ce, cz, an, dm = pkg.synth.head_candidates(K, seed=6)

    def carve(a, width):
        buf = torch.zeros(512 * width, dtype=torch.from_numpy(a).dtype, device="cuda")
        buf[: K * width] = dev(a).flatten()
        return buf[: K * width]
    ins = [carve(sc, 1).view(1, K), carve(cl, 1).view(1, K), carve(xs, 1).view(1, K), carve(ys, 1).view(1, K),
           carve(ce, 2).view(1, 1, K, 2), carve(cz, 1).view(1, 1, K, 1), carve(an, 1).view(1, 1, K, 1),
           carve(dm, 3).view(1, 1, K, 3)]
    fargs = (K, cfg.x_min, cfg.x_max, cfg.y_min, cfg.y_max, cfg.z_min, cfg.z_max, cfg.voxel_x, cfg.voxel_y,
             cfg.voxel_z, cfg.score_threshold)
    bm, vm = plg.add_filter_box_by_score_op(ours, *fargs).enqueue(ins, poison=float("nan"))
    br, vr = plg.add_filter_box_by_score_op(ref_lib(plg, "filterBoxByScore"), *fargs).enqueue(ins)
    assert int(vm[0]) == int(vr[0]) > 0
    n = int(vm[0])
    a, b = bm[0, :n].cpu().numpy(), br[0, :n].cpu().numpy()
    key = lambda t: t[np.lexsort(t.T[::-1])]
    assert np.array_equal(key(a), key(b))       # same boxes, bit exact, order canonicalised


def test_gather_scatter_vs_reference_kernels(plg, ours, cfgs, frame0):
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V = o["pillar_num"]
    owp = cpu.window_partition(o["coords"], V, cfg, 0)
    ogs = cpu.get_set(owp["global_index"], owp["coors_in_win"], owp["voxel_num_in_win"], owp["win_num"], cfg, 0)
    rng = np.random.default_rng(3)
    x = np.zeros((1, cfg.max_pillars_num, 192), np.float32)
    pos = np.zeros_like(x)
    x[0, :V] = rng.standard_normal((V, 192))
    pos[0, :V] = rng.standard_normal((V, 192))
    idx, ns = dev(ogs["global_index_in_set"])[None], i32(ogs["set_num"])
    for axis in (0, 1):
        a = (cfg.max_win_num, cfg.voxel_num_set, 192, axis)
        gm = plg.add_get_value_by_index_op(ours, *a).enqueue([dev(x), dev(pos), idx, ns], poison=float("nan"))
        gr = plg.add_get_value_by_index_op(ref_lib(plg, "getValueByIndex"), *a).enqueue([dev(x), dev(pos), idx, ns])
        for t_m, t_r in zip(gm, gr):
            assert torch.equal(t_m, t_r)
        feat = gm[2]     # any [1,800,36,192] tensor
        sm = plg.add_map_set_feature2voxel_op(ours, *a, cfg.max_pillars_num).enqueue([feat, idx, ns], poison=float("nan"))[0]
        sr = plg.add_map_set_feature2voxel_op(ref_lib(plg, "mapSetFeature2voxel"), *a, cfg.max_pillars_num).enqueue([feat, idx, ns])[0]
        assert torch.equal(sm, sr)


def test_scatter_max_and_map2bev_vs_reference_kernels(plg, ours, cfgs, frame0):
    """TorchScatterMaxPlugin / Map2BevPlugin against the reference's own kernels, created from the same fields."""
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V, Pc = o["pillar_num"], o["point_num"]
    rng = np.random.default_rng(9)
    for F in (96, 192):
        feat = np.zeros((1, cfg.max_points_num_voxel_filter, F), np.float32)
        feat[0, :Pc] = rng.standard_normal((Pc, F)) * 2
        ins = [dev(feat), dev(o["point_index_in_voxel"])[None], dev(o["point_num_in_voxel"].reshape(1, -1, 1)), i32(V)]
        a = (cfg.max_points_num_voxel_filter, cfg.max_pillars_num, F)
        mine = plg.add_torch_scatter_max(ours, *a).enqueue(ins, poison=float("nan"))
        ref = plg.add_torch_scatter_max(ref_lib(plg, "torchScatterMax"), *a).enqueue(ins)
        assert torch.equal(mine[0], ref[0]) and torch.equal(mine[1], ref[1])
        # optional 5th input (the voxeliser's row count): same result without the full clear
        mine5 = plg.add_torch_scatter_max(ours, *a).enqueue(ins + [i32(Pc)], poison=float("nan"))
        assert torch.equal(mine5[0], ref[0]) and torch.equal(mine5[1], ref[1])
    x = np.zeros((1, cfg.max_pillars_num, 192), np.float32)
    x[0, :V] = rng.standard_normal((V, 192))
    ins = [dev(x), dev(o["coords"])[None], i32(V)]
    a = (cfg.max_pillars_num, 192, cfg.grid_x, cfg.grid_y)
    mine = plg.add_map_2_bev_op(ours, *a).enqueue(ins, poison=float("nan"))[0]
    ref = plg.add_map_2_bev_op(ref_lib(plg, "map2bev"), *a).enqueue(ins)[0]
    assert tuple(mine.shape) == (1, cfg.grid_x, cfg.grid_y, 192) and torch.equal(mine, ref)
