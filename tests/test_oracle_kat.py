"""Pins the CPU oracle (oracle/dsvt_oracle.c) against the known-answer table derived from the reference's
own sample frames (tests/golden/kat.json) and against independent restatements."""
import os

import numpy as np
import pytest

from conftest import pad_points
from oracle import cpu

REF_BIN = "/root/reference/data/bin"


def _chain(points, cfg):
    o = cpu.points2features(pad_points(points, cfg.max_points_num), len(points), cfg)
    res = {"p2f": o}
    for which in (0, 1):
        wp = cpu.window_partition(o["coords"], o["pillar_num"], cfg, which)
        gs = cpu.get_set(wp["global_index"], wp["coors_in_win"], wp["voxel_num_in_win"], wp["win_num"], cfg, which)
        res[which] = (wp, gs)
    return res


def _check_kat(res, row):
    o = res["p2f"]
    assert o["pillar_num"] == row["pillars"]
    assert o["point_num"] == row["kept_points"]
    V = o["pillar_num"]
    assert int(o["point_num_in_voxel"][:V].sum()) == row["kept_points"]
    assert int((o["point_num_in_voxel"][:V] == 48).sum()) >= row["overfull_pillars"]
    for which, tag in ((0, "win12"), (1, "win24_shift6")):
        wp, gs = res[which]
        assert wp["win_num"] == row[tag]["windows"]
        assert int(wp["voxel_num_in_win"].max()) == row[tag]["max_voxels_per_window"]
        assert gs["set_num"] == row[tag]["sets"]
        ns = gs["set_num"]
        for plane in (0, 1):
            assert int((gs["set_voxel_mask"][plane, :ns] < 0).sum()) == row[tag]["masked_slots"]


def test_frame0_known_answers(frame0, cfgs, kat):
    assert len(frame0) == kat["000000"]["points"] == 34537
    res = _chain(frame0, cfgs.REFERENCE)
    _check_kat(res, kat["000000"])
    # the constants quoted in the reference's own source comments
    assert res["p2f"]["pillar_num"] == 5504 and res[0][1]["set_num"] == 454


@pytest.mark.skipif(not os.path.isdir(REF_BIN), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("name", ["000000", "000003", "000004"])
def test_all_reference_frames(name, cfgs, kat):
    pts = np.fromfile(os.path.join(REF_BIN, name + ".bin"), dtype=np.float32).reshape(-1, 4)
    _check_kat(_chain(pts, cfgs.REFERENCE), kat[name])


def test_voxeliser_structure(frame0, cfgs):
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V, P = o["pillar_num"], o["point_num"]
    coords = o["coords"][:V]
    cell = coords[:, 2].astype(np.int64) * cfg.grid_x + coords[:, 3]
    assert np.all(np.diff(cell) > 0), "canonical pillar order is ascending y*gx+x"
    assert np.all(coords[:, :2] == 0)
    n = o["point_num_in_voxel"][:V]
    rows = o["point_index_in_voxel"][:V]
    base = np.concatenate([[0], np.cumsum(n)[:-1]])
    for s in range(cfg.max_num_points_per_voxel):
        sel = n > s
        assert np.all(rows[sel, s] == base[sel] + s), "rows are pillar-major, slot-minor"
        assert np.all(rows[~sel, s] == 0)
    f = o["point_features"]
    assert np.all(f[P:] == 0) and np.all(o["coords"][V:] == 0) and np.all(o["point_num_in_voxel"][V:] == 0)
    # every kept point lies inside its pillar and channels 7..9 are offsets to the pillar centre
    pid = np.repeat(np.arange(V), n)
    ix = np.floor((f[:P, 0] - np.float32(cfg.x_min)) / np.float32(cfg.voxel_x)).astype(np.int64)
    iy = np.floor((f[:P, 1] - np.float32(cfg.y_min)) / np.float32(cfg.voxel_y)).astype(np.int64)
    assert np.all(ix == coords[pid, 3]) and np.all(iy == coords[pid, 2])
    cx = (ix + 0.5) * np.float64(np.float32(cfg.voxel_x)) + np.float64(np.float32(cfg.x_min))
    assert np.allclose(f[:P, 7], (f[:P, 0].astype(np.float64) - cx).astype(np.float32), atol=0, rtol=0)
    # channels 4..6: offset to the mean of the kept points
    mean_x = np.add.reduceat(f[:P, 0].astype(np.float64), base) / n
    assert np.allclose(f[:P, 4], f[:P, 0] - mean_x[pid].astype(np.float32), atol=2e-5)
    # overfull pillars keep the LOWEST input indices (serial order)
    pts = frame0
    inr = ((pts[:, 0] >= np.float32(cfg.x_min)) & (pts[:, 0] < np.float32(cfg.x_max)) &
           (pts[:, 1] >= np.float32(cfg.y_min)) & (pts[:, 1] < np.float32(cfg.y_max)) &
           (pts[:, 2] >= np.float32(cfg.z_min)) & (pts[:, 2] < np.float32(cfg.z_max)))
    pcx = np.floor((pts[:, 0] - np.float32(cfg.x_min)) / np.float32(cfg.voxel_x)).astype(np.int64)
    pcy = np.floor((pts[:, 1] - np.float32(cfg.y_min)) / np.float32(cfg.voxel_y)).astype(np.int64)
    pcell = np.where(inr, pcy * cfg.grid_x + pcx, -1)
    big = int(np.argmax(n == 48))
    members = np.nonzero(pcell == cell[big])[0]
    assert len(members) > 48
    got = f[base[big]: base[big] + 48, :4]
    assert np.array_equal(got, pts[members[:48]])


def test_get_set_formula(frame0, cfgs):
    """Independent restatement of DSVT eq.(3) and of the two orderings (getSet.cu:346, :388, :463)."""
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    for which in (0, 1):
        wx, wy, wz = cfg.win_shapes[which]
        wp = cpu.window_partition(o["coords"], o["pillar_num"], cfg, which)
        gs = cpu.get_set(wp["global_index"], wp["coors_in_win"], wp["voxel_num_in_win"], wp["win_num"], cfg, which)
        S, set_id = cfg.voxel_num_set, 0
        for w in range(wp["win_num"]):
            N = int(wp["voxel_num_in_win"][w])
            gi = wp["global_index"][w, :N]
            assert np.all(np.diff(gi) > 0), "canonical in-window order is ascending voxel id"
            c = wp["coors_in_win"][w, :N]
            ky = c[:, 1] * wx * wz + c[:, 2] * wz + c[:, 0]
            kx = c[:, 2] * wy * wz + c[:, 1] * wz + c[:, 0]
            assert len(np.unique(ky)) == N
            sy, sx = gi[np.argsort(ky, kind="stable")], gi[np.argsort(kx, kind="stable")]
            ns = -(-N // S)
            for j in range(ns):
                r = ((j * S + np.arange(S)) * N // S) // ns
                assert np.array_equal(gs["global_index_in_set"][0, set_id], sy[r])
                assert np.array_equal(gs["global_index_in_set"][1, set_id], sx[r])
                m = np.zeros(S, np.float32)
                m[1:][r[1:] == r[:-1]] = -np.finfo(np.float32).max
                assert np.array_equal(gs["set_voxel_mask"][0, set_id], m)
                assert np.array_equal(gs["mask_expand_0"][set_id], np.broadcast_to(m, (cfg.num_heads, S)))
                assert np.array_equal(gs["mask_expand_1"][set_id], gs["mask_expand_0"][set_id])
                set_id += 1
        assert set_id == gs["set_num"]
        ns_all = gs["set_num"]
        assert np.all(gs["global_index_in_set"][:, ns_all:] == 0) and np.all(gs["mask_expand_0"][ns_all:] == 0)
        # every voxel of the frame appears in exactly one set per ordering
        for plane in (0, 1):
            seen = np.unique(gs["global_index_in_set"][plane, :ns_all])
            assert np.array_equal(seen, np.arange(o["pillar_num"]))
        bits = gs["set_voxel_mask"].view(np.uint32)
        assert set(np.unique(bits).tolist()) <= {0x00000000, 0xFF7FFFFF}, "mask constants are 0 and -FLT_MAX"


def test_gelu_layernorm_filterbox_against_numpy(cfgs, pkg):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, 384)).astype(np.float32) * 3
    g = cpu.gelu(x, 50)
    xd = x[:50].astype(np.float64)
    ref = ((0.5 + 0.5 * np.tanh(xd * (0.035677408136300125 * xd * xd + 0.7978845608028654))) * xd).astype(np.float32)
    assert np.allclose(g[:50], ref, rtol=1e-6, atol=1e-7) and np.all(g[50:] == 0)   # libm vs numpy tanh: <= 1 ulp

    x = rng.standard_normal((40, 192)).astype(np.float32) * 2 + 1
    gamma, beta = rng.standard_normal(192).astype(np.float32), rng.standard_normal(192).astype(np.float32)
    y = cpu.layer_norm(x, 33, gamma, beta, eps=0.0)
    xd = x[:33].astype(np.float64)
    mu, var = xd.mean(1, keepdims=True), xd.var(1, keepdims=True)
    assert np.allclose(y[:33], (xd - mu) / np.sqrt(var) * gamma + beta, atol=2e-5) and np.all(y[33:] == 0)
    res = rng.standard_normal((40, 192)).astype(np.float32)
    assert np.array_equal(cpu.layer_norm(x, 33, gamma, beta, residual=res)[:33],
                          cpu.layer_norm(x + res, 33, gamma, beta)[:33])

    cfg = cfgs.REFERENCE
    sc, cl, xs, ys, ce, cz, an, dm = pkg.synth.head_candidates(cfg.max_top_k, seed=3)
    boxes, valid, kept = cpu.filter_box(sc, cl, xs, ys, ce, cz, an, dm, cfg)
    nx = (xs.astype(np.float32) + ce[:, 0]) * np.float32(cfg.voxel_x) + np.float32(cfg.x_min)
    ny = (ys.astype(np.float32) + ce[:, 1]) * np.float32(cfg.voxel_y) + np.float32(cfg.y_min)
    keep = ((nx >= cfg.x_min) & (nx < cfg.x_max) & (ny >= cfg.y_min) & (ny < cfg.y_max) &
            (cz >= cfg.z_min) & (cz < cfg.z_max) & (sc >= np.float32(cfg.score_threshold)))
    # candidates within one ulp of a range edge may flip between the FMA and the two-step evaluation
    edge = (np.abs(nx - cfg.x_max) < 1e-4) | (np.abs(ny - cfg.y_max) < 1e-4)
    assert 0 < valid < cfg.max_top_k
    assert set(kept.tolist()) ^ set(np.nonzero(keep)[0].tolist()) <= set(np.nonzero(edge)[0].tolist())
    assert np.all(np.diff(kept) > 0), "canonical output order is ascending candidate index"
    assert np.array_equal(boxes[:valid, 8], sc[kept]) and np.array_equal(boxes[:valid, 7], cl[kept].astype(np.float32))
    assert np.all(boxes[valid:] == 0)


def test_attention_oracle_matches_torch_mha(attention_case):
    """multHeadAttention() restates nn.MultiheadAttention; the fixture output comes from
    torch.nn.functional.multi_head_attention_forward (tools/make_golden.py)."""
    c = attention_case
    out = cpu.set_attention(c["q"], c["k"], c["v"], c["mask"], c["q"].shape[0], c["w_in"], c["b_in"], c["w_out"],
                            c["b_out"])
    assert np.abs(out - c["out"]).max() < 2e-5


def test_scatter_max_and_map2bev_against_numpy(frame0, cfgs):
    """Oracle restatements of torchScatterMax.cu:201-262 and map2bev.cu:250-265 against independent numpy code."""
    cfg = cfgs.REFERENCE
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V, Pc = o["pillar_num"], o["point_num"]
    rng = np.random.default_rng(0)
    F = 96
    feat = np.zeros((cfg.max_points_num_voxel_filter, F), np.float32)
    feat[:Pc] = rng.standard_normal((Pc, F)) * 2
    feat[0] = -3000000.0                       # below the -1000000 start value: the max of a 1-point pillar is then -1000000
    mp, mv = cpu.torch_scatter_max(feat, o["point_index_in_voxel"], o["point_num_in_voxel"], V)
    piv, pnv = o["point_index_in_voxel"], o["point_num_in_voxel"].reshape(-1)
    for v in list(range(0, V, 97)) + [V - 1]:
        rows = piv[v, : pnv[v]]
        want = np.maximum(feat[rows].max(axis=0), np.float32(-1000000.0))
        assert np.array_equal(mv[v], want)
        assert np.array_equal(mp[rows], np.broadcast_to(want, (len(rows), F)))
    assert np.all(mv[V:] == 0) and np.all(mp[Pc:] == 0)
    # every kept point row is covered by exactly one pillar
    assert sorted(np.concatenate([piv[v, : pnv[v]] for v in range(V)]).tolist()) == list(range(Pc))

    x = np.zeros((cfg.max_pillars_num, 192), np.float32)
    x[:V] = rng.standard_normal((V, 192))
    bev = cpu.map2bev(x, o["coords"], V, cfg.grid_x, cfg.grid_y)
    want = np.zeros((cfg.grid_y, cfg.grid_x, 192), np.float32)
    want[o["coords"][:V, 2], o["coords"][:V, 3]] = x[:V]
    assert np.array_equal(bev, want)
