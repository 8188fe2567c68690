"""GPU parity of a whole DSVT block (src/dsvt-ai-trt.cpp:653-756) as the bench frame runs it:
voxeliser -> window partition -> getSet -> 2 x [fused set attention -> LayerNorm(y + x) -> FC 192->384 -> GELU ->
FC 384->192 -> LayerNorm(src + src2) -> LayerNorm(src + x)] -> block LayerNorm, against the CPU oracle chained the same
way (the FFN linears, TensorRT FullyConnected layers in the reference, are restated in float64 numpy).

Tolerance: the north-star's FP32 bound is 1e-3 abs; the chain is held to 2e-4 (every stage ends in a LayerNorm, so
errors do not grow along the block)."""
import importlib

import numpy as np
import pytest
import torch

from conftest import pad_points
from oracle import cpu

pytestmark = pytest.mark.gpu


def oracle_block(fr, w, cfg, frame0):
    """The block of HotPathFrame.run() (block 0: 12x12 windows) with the oracle's functions."""
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V, mp = o["pillar_num"], cfg.max_pillars_num
    owp = cpu.window_partition(o["coords"], V, cfg, 0)
    ogs = cpu.get_set(owp["global_index"], owp["coors_in_win"], owp["voxel_num_in_win"], owp["win_num"], cfg, 0)
    ns, idx, mask = ogs["set_num"], ogs["global_index_in_set"], ogs["mask_expand_0"]
    gamma, beta = w.gamma.cpu().numpy(), w.beta.cpu().numpy()
    x = fr.x0.cpu().numpy().copy()
    x[V:] = 0                     # rows >= voxel_num are never read by the plugins
    x_in, ln = x, 0
    for enc in (0, 1):
        pos = fr.pos[0][enc].cpu().numpy()
        q, k, v = cpu.get_value_by_index(x, pos, idx, ns, enc)
        a = cpu.set_attention(q, k, v, mask, ns, *w.attn_host[enc])
        y = cpu.map_set_feature2voxel(a, idx, ns, enc, mp)
        src = cpu.layer_norm(y, V, gamma[ln], beta[ln], residual=x); ln += 1
        w1, b1, w2, b2 = w._ffn_host[enc]
        h = np.zeros((mp, cfg.ffn_channel_num), np.float32)
        h[:V] = (src[:V].astype(np.float64) @ w1.T.astype(np.float64) + b1).astype(np.float32)
        g = cpu.gelu(h, V)
        src2 = np.zeros((mp, cfg.channel_num), np.float32)
        src2[:V] = (g[:V].astype(np.float64) @ w2.T.astype(np.float64) + b2).astype(np.float32)
        src = cpu.layer_norm(src, V, gamma[ln], beta[ln], residual=src2); ln += 1
        x = cpu.layer_norm(src, V, gamma[ln], beta[ln], residual=x); ln += 1
    return cpu.layer_norm(x, V, gamma[ln], beta[ln], residual=x_in), V


@pytest.mark.parametrize("ffn", ["graph", "fused", "epilogue"])
@pytest.mark.parametrize("fuse_ln", [True, False])
def test_dsvt_block_chain(frame0, cfgs, ffn, fuse_ln):
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    cfg = cfgs.REFERENCE.with_(num_blocks=1)
    w = pipeline.FrameWeights(cfg, seed=3)
    fr = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC, seed=5, ffn=ffn, fuse_ln=fuse_ln)
    fr.load_points(frame0)
    fr.run()
    torch.cuda.synchronize()
    ref, V = oracle_block(fr, w, cfg, frame0)
    assert int(fr.vox.pillar_num[0]) == V
    got = fr.final.cpu().numpy()
    assert np.all(got[V:] == 0)
    assert np.isfinite(got).all()
    err = np.abs(got - ref).max()
    assert err <= 2e-4, err
    # the frame really went through the C ABI (voxeliser, partition, plans, attention, LayerNorms, FFN linears)
    assert fr.launches_per_frame > 20


def _lin64(x, W, scale=None, shift=None, bias=None, relu=False):
    """float64 restatement of a TensorRT FullyConnected (+ Scale of the folded BatchNorm) (+ ReLU) layer."""
    y = x.astype(np.float64) @ W.T.astype(np.float64)
    if scale is not None:
        y = y * scale + shift
    if bias is not None:
        y = y + bias
    if relu:
        y = np.maximum(y, 0.0)
    return y.astype(np.float32)


def oracle_backbone(w, cfg, frame0):
    """The whole 3-D backbone of HotPathFrame(backbone=True) with the oracle's functions (src/dsvt-ai-trt.cpp:571-1128)."""
    o = cpu.points2features(pad_points(frame0, cfg.max_points_num), len(frame0), cfg)
    V, Pc, mp, Pm = o["pillar_num"], o["point_num"], cfg.max_pillars_num, cfg.max_points_num_voxel_filter
    piv, pnv = o["point_index_in_voxel"], o["point_num_in_voxel"]
    w0, s0, t0 = w.vfe_host["pfn0"]
    w1, s1, t1 = w.vfe_host["pfn1"]
    h0 = np.zeros((Pm, cfg.pfn_channels[0]), np.float32)
    h0[:Pc] = _lin64(o["point_features"][:Pc], w0, s0, t0, relu=True)                      # :577
    mp0, _ = cpu.torch_scatter_max(h0, piv, pnv, V)                                        # :579
    h1 = np.zeros((Pm, cfg.pfn_channels[1]), np.float32)
    h1[:Pc] = _lin64(np.concatenate([h0[:Pc], mp0[:Pc]], axis=1), w1, s1, t1, relu=True)   # :583-587
    _, x = cpu.torch_scatter_max(h1, piv, pnv, V)                                          # :589, output 1
    parts = []
    for which in (0, 1):
        owp = cpu.window_partition(o["coords"], V, cfg, which)
        ogs = cpu.get_set(owp["global_index"], owp["coors_in_win"], owp["voxel_num_in_win"], owp["win_num"], cfg, which)
        parts.append((owp, ogs))
    gamma, beta = w.gamma.cpu().numpy(), w.beta.cpu().numpy()
    ln = 0
    for blk in range(cfg.num_blocks):
        ogs = parts[blk % 2][1]
        ns, idx, mask = ogs["set_num"], ogs["global_index_in_set"], ogs["mask_expand_0"]
        x_in = x
        for enc in (0, 1):
            a, sc, sh, b2, bias2 = w.pos_host[blk][enc]
            pos = np.zeros((mp, cfg.channel_num), np.float32)
            xy = parts[enc][0]["coors_in_win_x_y"][:V]                                     # shift-`enc` window coordinates
            pos[:V] = _lin64(_lin64(xy, a, sc, sh, relu=True), b2, bias=bias2)             # :603-637
            q, k, v = cpu.get_value_by_index(x, pos, idx, ns, enc)
            att = cpu.set_attention(q, k, v, mask, ns, *w.attn_host[blk * 2 + enc])
            y = cpu.map_set_feature2voxel(att, idx, ns, enc, mp)
            src = cpu.layer_norm(y, V, gamma[ln], beta[ln], residual=x); ln += 1
            f1, fb1, f2, fb2 = w._ffn_host[blk * 2 + enc]
            h = np.zeros((mp, cfg.ffn_channel_num), np.float32)
            h[:V] = _lin64(src[:V], f1, bias=fb1)
            g = cpu.gelu(h, V)
            src2 = np.zeros((mp, cfg.channel_num), np.float32)
            src2[:V] = _lin64(g[:V], f2, bias=fb2)
            src = cpu.layer_norm(src, V, gamma[ln], beta[ln], residual=src2); ln += 1
            x = cpu.layer_norm(src, V, gamma[ln], beta[ln], residual=x); ln += 1
        x = cpu.layer_norm(x, V, gamma[ln], beta[ln], residual=x_in); ln += 1
    bev = cpu.map2bev(x, o["coords"], V, cfg.grid_x, cfg.grid_y)
    return x, bev, V, Pc


@pytest.mark.parametrize("ffn", ["graph", "fused", "epilogue", "kernel", "layer"])
def test_whole_3d_backbone_chain(frame0, cfgs, ffn):
    """Raw points -> PFN -> scatter-max -> partition / sets -> 2 DSVT blocks (12x12 and shifted 24x24 windows) -> BEV map:
    every layer of the reference's 3-D backbone executed on the GPU through the C ABI, against the oracle chain."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    cfg = cfgs.REFERENCE.with_(num_blocks=2)
    w = pipeline.FrameWeights(cfg, seed=4)
    fr = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC, seed=6, ffn=ffn, backbone=True)
    fr.load_points(frame0)
    fr.run()
    torch.cuda.synchronize()
    ref, bev, V, Pc = oracle_backbone(w, cfg, frame0)
    assert int(fr.vox.pillar_num[0]) == V and int(fr.vox.point_num[0]) == Pc
    # VFE output (pillar features before the blocks)
    vfe = fr.max_voxel[-1].cpu().numpy()
    assert np.all(vfe[V:] == 0)
    got = fr.final.cpu().numpy()
    assert np.isfinite(got).all() and np.all(got[V:] == 0)
    err = np.abs(got - ref).max()
    assert err <= 5e-4, err                     # north-star FP32 bound: 1e-3
    gbev = fr.bev.cpu().numpy()
    assert np.array_equal(gbev != 0, bev != 0)  # the same cells are occupied
    assert np.abs(gbev - bev).max() <= 5e-4


@pytest.mark.parametrize("precision,tol", [("DSVT_ATTN_FP32", 5e-4), ("DSVT_ATTN_FP16", 1e-2)])
def test_backbone_chain_single_kernel_precisions(frame0, cfgs, precision, tol):
    """The headline frame kind (norms in the epilogues) with the single-kernel set attention precisions -- the bench's
    fp16_config leg (BASELINE.json configs[2], tolerance 1e-2) and the CUDA-core FP32 path: dsvt_set_attention_fused_norm_launch
    falls back to attention + row-wise LayerNorm there."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    cfg = cfgs.REFERENCE.with_(num_blocks=2)
    w = pipeline.FrameWeights(cfg, seed=4)
    fr = _run_backbone(cfg, w, frame0, ffn="epilogue", precision=getattr(capi, precision))
    _check_backbone(fr, w, cfg, frame0, tol)


# ---- the bench workload and the trained weights -----------------------------------------
def _run_backbone(cfg, w, cloud, seed=6, ffn="graph", precision=None):
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    fr = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC if precision is None else precision, seed=seed,
                               ffn=ffn, backbone=True)
    fr.load_points(cloud)
    fr.run()
    torch.cuda.synchronize()
    return fr


def _check_backbone(fr, w, cfg, cloud, tol):
    ref, bev, V, Pc = oracle_backbone(w, cfg, cloud)
    assert int(fr.vox.pillar_num[0]) == V and int(fr.vox.point_num[0]) == Pc
    got = fr.final.cpu().numpy()
    assert np.isfinite(got).all() and np.all(got[V:] == 0)
    err = np.abs(got - ref).max()
    assert err <= tol, err
    gbev = fr.bev.cpu().numpy()
    assert np.array_equal(gbev != 0, bev != 0)
    assert np.abs(gbev - bev).max() <= tol
    return V, err


@pytest.mark.parametrize("geometry", ["WAYMO", "WAYMO_030"])
def test_bench_workload_chain(pkg, cfgs, geometry):
    """The bench's own frame (BASELINE.json configs[1]): one 200k-point ring cloud, Waymo capacities (320k / 40k / 4096),
    ALL 4 blocks, every layer of the 3-D backbone, pillar 0.32 (grid 468) and 0.30 (grid 500), against the oracle chain."""
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    cfg = getattr(cfgs, geometry)
    cloud = pkg.synth.ring_lidar(200000, seed=0)
    w = pipeline.FrameWeights(cfg, seed=0)
    fr = _run_backbone(cfg, w, cloud, ffn="layer")          # the bench's frame kind (bench.FRAME_KINDS["backbone3d"])
    V, err = _check_backbone(fr, w, cfg, cloud, 5e-4)
    assert V > 25000                    # the survey's density (~30.6 k pillars), not round 1's 16.6 k
    assert int(fr.gs[0].set_num[0]) > 1300 and int(fr.gs[1].set_num[0]) > 900


def test_trained_weights_chain(frame0, cfgs):
    """BASELINE.json configs[0]: the reference's sample frame through all 4 blocks with the TRAINED tensors of dsvt.wts
    (tests/golden/dsvt_backbone3d_wts.npz, written by tools/make_wts_fixture.py through oracle/wts.py; in_proj split as
    helper.h:367-433; in_proj_bias is all zeros in the real file)."""
    import os
    from conftest import GOLDEN
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    t = dict(np.load(os.path.join(GOLDEN, "dsvt_backbone3d_wts.npz")))
    cfg = cfgs.REFERENCE
    w = pipeline.FrameWeights.from_wts(cfg, t)
    assert all(float(np.abs(a[1]).max()) == 0.0 for a in w.attn_host)      # the edge case is really exercised
    for ffn in ("graph", "fused", "epilogue"):
        fr = _run_backbone(cfg, w, frame0, ffn=ffn)
        V, err = _check_backbone(fr, w, cfg, frame0, 5e-4)
        assert V == 5504


@pytest.mark.parametrize("n_points,seed", [(200000, 0), (60000, 3), (3000, 1)])
def test_vfe_fused_equals_separate_launches(pkg, cfgs, n_points, seed):
    """dsvt_vfe_fused_launch (PFN 0 + scatter-max + concat + PFN 1 + scatter-max in one kernel, no per-point tensor in memory)
    against the four separate launches on the voxeliser's real output: same arithmetic in the same order, bit for bit."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    cfg = cfgs.WAYMO
    w = pipeline.FrameWeights(cfg, seed=2)
    cloud = pkg.synth.ring_lidar(n_points, seed=seed)
    sep = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC, seed=1, ffn="epilogue", backbone=True)
    sep.load_points(cloud)
    sep.run()
    one = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC, seed=1, ffn="kernel", backbone=True)
    one.load_points(cloud)
    one.max_voxel[-1].fill_(float("nan"))
    one.run()
    torch.cuda.synchronize()
    V = int(sep.vox.pillar_num[0])
    assert V == int(one.vox.pillar_num[0]) and V > 0
    a, b = sep.max_voxel[-1], one.max_voxel[-1]
    assert torch.isfinite(b).all() and torch.all(b[V:] == 0)
    assert torch.equal(a[:V], b[:V])


@pytest.mark.parametrize("n_points", [0, 1, 47])
def test_headline_frame_degenerate_clouds(pkg, cfgs, n_points):
    """The bench's frame kind on an empty cloud, a single point and a single pillar's worth of points: no fault, counts right,
    every output row beyond the pillar count zero, and (for n > 0) finite features equal to the separate-launch form's."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    cfg = cfgs.REFERENCE
    w = pipeline.FrameWeights(cfg, seed=9)
    cloud = np.zeros((n_points, 4), np.float32)
    if n_points:
        rng = np.random.default_rng(n_points)
        cloud[:, :2] = 10.0 + (rng.random((n_points, 2)) * (0.2 if n_points > 1 else 0.0)).astype(np.float32)   # one pillar
        cloud[:, 2] = -1.0
        cloud[:, 3] = 0.5
    frames = {}
    for ffn in ("layer", "epilogue"):
        fr = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC, seed=2, ffn=ffn, backbone=True)
        fr.load_points(cloud)
        fr.run()
        torch.cuda.synchronize()
        frames[ffn] = fr
    one, sep = frames["layer"], frames["epilogue"]
    V = int(one.vox.pillar_num[0])
    assert V == int(sep.vox.pillar_num[0]) and V == (1 if n_points else 0)
    assert int(one.vox.point_num[0]) == n_points
    assert bool((one.max_voxel[-1][V:] == 0).all()) and bool((one.final[V:] == 0).all())
    if V:
        assert torch.isfinite(one.final[:V]).all()
        assert torch.equal(one.max_voxel[-1][:V], sep.max_voxel[-1][:V])
        assert (one.final[:V] - sep.final[:V]).abs().max().item() <= 1e-5
    assert bool((one.bev.reshape(-1, cfg.channel_num).abs().sum(1) > 0).sum() == V) or V == 0


def test_layer_tail_kernel_equals_four_kernel_layer(frame0, cfgs):
    """dsvt_attention_tail_ffn_launch (out-projection + norm1 + FFN + norms in one kernel) against the form it replaces
    (dsvt_set_attention_fused_norm_launch + dsvt_ffn_fused_launch) over two whole blocks: same arithmetic, bit for bit."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    cfg = cfgs.REFERENCE.with_(num_blocks=2)
    w = pipeline.FrameWeights(cfg, seed=4)
    a = _run_backbone(cfg, w, frame0, ffn="kernel")
    b = _run_backbone(cfg, w, frame0, ffn="layer")
    V = int(a.vox.pillar_num[0])
    assert V > 1000 and torch.equal(a.final[:V], b.final[:V]) and bool((b.final[V:] == 0).all())
    assert b.launches_per_frame < a.launches_per_frame
