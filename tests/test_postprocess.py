"""CenterHead post-process graph + rotated NMS (SURVEY 8(f) #4 tail): the oracle against the reference's own host NMS
(CPU), and the CUDA kernels against the oracle (GPU)."""
import ctypes
import importlib
import os

import numpy as np
import pytest

from oracle import build as obuild
from oracle import cpu


def synth_boxes(n, seed, spread=30.0):
    rng = np.random.default_rng(seed)
    b = np.zeros((n, 9), np.float32)
    b[:, 0:2] = rng.uniform(-spread, spread, (n, 2))
    b[:, 2] = rng.uniform(-2, 1, n)
    b[:, 3:6] = np.exp(rng.normal(0.8, 0.4, (n, 3)))
    b[:, 6] = rng.uniform(-1.57, 1.57, n)
    b[:, 7] = rng.integers(0, 10, n)
    b[:, 8] = rng.uniform(0.3, 1.0, n)
    return b


def synth_head(nc, H, W, seed):
    rng = np.random.default_rng(seed)
    hm = (rng.standard_normal((nc, H, W)) * 1.5 - 3.0).astype(np.float32)
    return (hm, rng.uniform(0, 1, (2, H, W)).astype(np.float32), rng.uniform(-3, 1, (1, H, W)).astype(np.float32),
            rng.normal(0.5, 0.4, (3, H, W)).astype(np.float32), rng.uniform(-1, 1, (2, H, W)).astype(np.float32))


@pytest.mark.parametrize("thresh", [0.01, 0.3])
def test_oracle_nms_equals_reference_host_nms(thresh):
    """oracle/_ref/libref_nms.so = the reference's include/helper.h compiled unmodified (oracle/build.py)."""
    so = obuild.build_reference_nms()
    if so is None:
        pytest.skip("oracle/_ref/libref_nms.so not built (no /root/reference here)")
    ref = ctypes.CDLL(so)
    for n, seed, spread in ((300, 0, 30.0), (500, 1, 12.0), (1, 2, 5.0), (37, 3, 3.0)):
        boxes = synth_boxes(n, seed, spread)
        k1 = np.zeros(n, np.int32)
        n1 = ref.ref_nms_cpu(boxes.ctypes.data_as(ctypes.c_void_p), n, ctypes.c_float(thresh), k1.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(cpu.nms(boxes, n, thresh), k1[:n1])
    assert len(cpu.nms(synth_boxes(5, 0), 0, thresh)) == 0


def test_oracle_topk_two_stage_is_global_topk():
    hm, ce, cz, dm, ro = synth_head(10, 60, 52, 4)
    o = cpu.center_head_topk(hm, ce, cz, dm, ro, K=500)
    flat = hm.reshape(-1)
    order = np.argsort(-flat, kind="stable")[:500]                  # logit descending, flat index ascending
    assert np.array_equal(o["classes"], (order // (60 * 52)).astype(np.int32))
    assert np.array_equal(o["ys"] * 52 + o["xs"], (order % (60 * 52)).astype(np.int32))
    assert np.all(np.diff(o["scores"]) <= 0) and o["scores"].dtype == np.float32
    assert np.allclose(o["dim"], np.exp(dm.reshape(3, -1)[:, order % (60 * 52)].T)) and o["dim"].shape == (500, 3)


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(10, 468, 468), (10, 60, 52), (3, 17, 31)])
def test_center_head_topk_gpu(shape):
    import torch
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    nc, H, W = shape
    K = 500
    hm, ce, cz, dm, ro = synth_head(nc, H, W, nc + H)
    hm.reshape(-1)[7] = hm.reshape(-1)[11] = 9.0                   # an exact tie at the top: ascending flat index wins
    op = capi.CenterHeadTopK(nc, H, W, K)
    d = lambda a: torch.from_numpy(a).cuda()[None]
    op(d(hm), d(ce), d(cz), d(dm), d(ro))
    torch.cuda.synchronize()
    o = cpu.center_head_topk(hm, ce, cz, dm, ro, K)
    assert np.array_equal(op.classes[0].cpu().numpy(), o["classes"])
    assert np.array_equal(op.xs[0].cpu().numpy(), o["xs"]) and np.array_equal(op.ys[0].cpu().numpy(), o["ys"])
    assert np.abs(op.scores[0].cpu().numpy() - o["scores"]).max() <= 1e-6
    assert np.array_equal(op.center[0, 0].cpu().numpy(), o["center"]) and np.array_equal(op.center_z[0, 0, :, 0].cpu().numpy(), o["center_z"])
    assert np.abs(op.dim[0, 0].cpu().numpy() - o["dim"]).max() <= 1e-5 * np.abs(o["dim"]).max()
    assert np.abs(op.angle[0, 0, :, 0].cpu().numpy() - o["angle"]).max() <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,spread,thresh", [(300, 0, 30.0, 0.01), (500, 1, 12.0, 0.01), (500, 5, 12.0, 0.3), (1, 2, 5.0, 0.01),
                                                  (0, 2, 5.0, 0.01), (37, 3, 3.0, 0.1)])
def test_rotated_nms_gpu(n, seed, spread, thresh):
    import torch
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    K = 500
    boxes = np.zeros((K, 9), np.float32)
    boxes[:n] = synth_boxes(max(n, 1), seed, spread)[:n]
    op = capi.RotatedNms(K, thresh)
    op.boxes.fill_(float("nan"))
    op(torch.from_numpy(boxes).cuda()[None], torch.tensor([n], dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    keep = cpu.nms(boxes, n, thresh)
    got_n = int(op.num[0])
    got_keep = op.keep[0, :got_n].cpu().numpy()
    if not np.array_equal(got_keep, keep):
        # a decision may only differ where the IoU sits within float noise of the threshold (cosf / sinf / atan2f last bits)
        margins = [abs(cpu.nms_iou(boxes, i, j) - thresh) for i in range(n) for j in range(i + 1, n)
                   if abs(cpu.nms_iou(boxes, i, j) - thresh) < 1e-5]
        assert margins, (got_n, len(keep))
    else:
        out = op.boxes[0].cpu().numpy()
        assert np.array_equal(out[:got_n], boxes[keep]) and np.all(out[got_n:] == 0)


@pytest.mark.gpu
def test_whole_pipeline_postprocess_parity():
    """The whole_pipeline frame kind (points -> 3-D backbone -> BEV map -> cuDNN stand-in convolutions -> post-process graph ->
    FilterBoxByScore -> rotated NMS): the post-process half is checked against the oracle on the head maps the frame really
    produced.  Those maps come out of BF16 convolutions, so the heat map is full of exact ties -- the tie order (ascending
    flat index) has to hold for the candidate lists to match."""
    import torch
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    synth = importlib.import_module("dsvt-ai-trt_b200.synth")
    config = importlib.import_module("dsvt-ai-trt_b200.config")
    cfg = config.WAYMO
    w = pipeline.FrameWeights(cfg, seed=3)
    f = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC, seed=3, ffn="epilogue", backbone=True, head="conv")
    f.load_points(synth.ring_lidar(60_000, seed=5))
    f.calibrate_head(n_above=180)
    f.run()
    torch.cuda.synchronize()
    m = {k: v[0].cpu().numpy() for k, v in f.conv(f.bev).items()}
    n_ties = m["hm"].size - np.unique(m["hm"]).size
    assert n_ties > 1000                                            # the premise of the test
    o = cpu.center_head_topk(m["hm"], m["center"], m["center_z"], m["dim"], m["rot"], cfg.max_top_k)
    t = f.topk
    assert np.array_equal(t.classes[0].cpu().numpy(), o["classes"])
    assert np.array_equal(t.xs[0].cpu().numpy(), o["xs"]) and np.array_equal(t.ys[0].cpu().numpy(), o["ys"])
    # FilterBoxByScore + NMS on the GPU's own candidates (last-bit differences of exp / atan must not be blamed on them)
    g = lambda a: a.cpu().numpy()
    boxes, valid, _ = cpu.filter_box(g(t.scores[0]), g(t.classes[0]), g(t.xs[0]), g(t.ys[0]), g(t.center[0, 0]), g(t.center_z[0, 0, :, 0]),
                                     g(t.angle[0, 0, :, 0]), g(t.dim[0, 0]), cfg)
    assert int(f.valid[0]) == valid and 50 <= valid <= cfg.max_top_k
    assert np.array_equal(g(f.boxes[0]), boxes)
    keep = cpu.nms(boxes, valid, 0.01)
    n = int(f.nms.num[0])
    assert n == len(keep) and np.array_equal(g(f.nms.keep[0, :n]), keep)
    assert np.array_equal(g(f.nms.boxes[0, :n]), boxes[keep])
