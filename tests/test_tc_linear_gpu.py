"""tcgen05 dense linear layer (dsvt_linear_*) against a float64 torch reference."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,tol", [(1, 2e-3), (2, 2e-3)])     # TF32 / FP16 operands, FP32 accumulate
@pytest.mark.parametrize("M,N,K", [(128, 64, 192), (300, 192, 192), (1000, 576, 192), (77, 384, 192), (5000, 192, 256)])
def test_tc_linear(precision, tol, M, N, K):
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    rng = np.random.default_rng(M + N)
    x = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) * 0.06).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    lin = capi.Linear(W, b, precision=precision)
    y = lin(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    ref = x.astype(np.float64) @ W.astype(np.float64).T + b
    err = np.abs(y.cpu().numpy() - ref).max()
    # 11-bit significands on both operands, K-long FP32 accumulation: error ~ sqrt(K) * 2^-11 * |x||w| ~ 5e-4
    assert err <= tol * (K / 192) ** 0.5, err


def _gelu_ref(v):       # gelu.cu:201-211 in double
    v = v.astype(np.float64)
    return ((0.5 + 0.5 * np.tanh(v * (0.035677408136300125 * v * v + 0.7978845608028654))) * v)


@pytest.mark.parametrize("N,K", [(192, 192), (384, 192), (192, 384), (576, 384)])
def test_linear_fp32_accurate_tensor_cores(N, K):
    """FP32_TC linear (FP16 hi+lo split operands, 3 tcgen05 MMAs per product) against a float64 matmul: the FFN
    (192->384->192) and PFN / pos-embed (192->192) shapes of src/dsvt-ai-trt.cpp:268-286,:461-529."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    rng = np.random.default_rng(N + K)
    M, cap = 1000, 1300
    W = (rng.standard_normal((N, K)) * 0.06).astype(np.float32)
    b = (rng.standard_normal(N) * 0.1).astype(np.float32)
    x = np.zeros((cap, K), np.float32)
    x[:M] = rng.standard_normal((M, K)) * 1.5
    ref = x[:M].astype(np.float64) @ W.astype(np.float64).T + b
    lin = capi.Linear(W, b, precision=capi.DSVT_ATTN_FP32_TC)
    y = lin(torch.from_numpy(x[:M]).cuda())
    assert np.abs(y.cpu().numpy() - ref).max() <= 2e-5
    rows = torch.tensor([M], dtype=torch.int32, device="cuda")
    out = torch.full((cap, N), float("nan"), device="cuda")
    lin.rows(torch.from_numpy(x).cuda(), rows, out=out)
    got = out.cpu().numpy()
    assert np.abs(got[:M] - ref).max() <= 2e-5 and np.all(got[M:] == 0)
    # fused GELU epilogue (the FFN's first linear + GeluPlugin in one pass)
    lin.rows(torch.from_numpy(x).cuda(), rows, activation=1, out=out)
    got = out.cpu().numpy()
    assert np.abs(got[:M] - _gelu_ref(ref)).max() <= 3e-5 and np.all(got[M:] == 0)
    # single FP16 operands: the USE_FP16 tolerance
    y16 = capi.Linear(W, b, precision=capi.DSVT_ATTN_FP16_GEMM)(torch.from_numpy(x[:M]).cuda())
    assert np.abs(y16.cpu().numpy() - ref).max() <= 1e-2


@pytest.mark.parametrize("N,K", [(96, 10), (192, 2), (96, 7), (192, 16)])
@pytest.mark.parametrize("rows", [0, 1, 1000, 1537])
def test_small_linear(N, K, rows):
    """Narrow first layers of the PFN / position-embedding MLPs (FullyConnected + Scale + ReLU, reference
    src/dsvt-ai-trt.cpp:268-286): y = relu((x W^T) * scale + shift) against float64; batch 2 with different row counts."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    rng = np.random.default_rng(N + K + rows)
    cap = 1537
    W = (rng.standard_normal((N, K)) * 0.3).astype(np.float32)
    scale = (0.5 + 0.1 * rng.standard_normal(N)).astype(np.float32)
    shift = (0.1 * rng.standard_normal(N)).astype(np.float32)
    x = (rng.standard_normal((2, cap, K)) * 3).astype(np.float32)
    counts = np.array([rows, min(cap, rows + 5)], np.int32)
    lin = capi.SmallLinear(W, scale, shift)
    out = torch.full((2, cap, N), float("nan"), device="cuda")
    lin(torch.from_numpy(x).cuda(), torch.from_numpy(counts).cuda(), activation=2, out=out)
    got = out.cpu().numpy()
    for b in range(2):
        ref = np.maximum((x[b, :counts[b]].astype(np.float64) @ W.T.astype(np.float64)) * scale + shift, 0.0)
        assert np.abs(got[b, :counts[b]] - ref).max(initial=0.0) <= 1e-5
        assert np.all(got[b, counts[b]:] == 0)
    # no activation, no BatchNorm
    lin2 = capi.SmallLinear(W)
    y2 = lin2(torch.from_numpy(x[0]).cuda(), torch.from_numpy(counts[:1]).cuda(), activation=0).cpu().numpy()
    assert np.abs(y2[:rows] - x[0, :rows].astype(np.float64) @ W.T.astype(np.float64)).max(initial=0.0) <= 1e-5


@pytest.mark.parametrize("rows", [0, 777, 2048])
@pytest.mark.parametrize("act", [0, 2])
def test_linear_rows_concat(rows, act):
    """PFN layer 1: Linear(192 -> 192) on the concatenation [x_lo (96) | x_hi (96)] read in place (reference
    src/dsvt-ai-trt.cpp:583-587), BatchNorm folded into W / b, ReLU epilogue; FP32 tolerance 2e-5 relative to the row scale."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    rng = np.random.default_rng(rows + act)
    cap = 2048
    W = (rng.standard_normal((192, 192)) * 0.07).astype(np.float32)
    b = (rng.standard_normal(192) * 0.1).astype(np.float32)
    lo = rng.standard_normal((cap, 96)).astype(np.float32)
    hi = rng.standard_normal((cap, 96)).astype(np.float32)
    lin = capi.Linear(W, b, precision=capi.DSVT_ATTN_FP32_TC)
    out = torch.full((cap, 192), float("nan"), device="cuda")
    lin.rows_concat(torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda(),
                    torch.tensor([rows], dtype=torch.int32, device="cuda"), activation=act, out=out)
    got = out.cpu().numpy()
    ref = np.concatenate([lo, hi], axis=1)[:rows].astype(np.float64) @ W.T.astype(np.float64) + b
    if act == 2:
        ref = np.maximum(ref, 0.0)
    assert np.abs(got[:rows] - ref).max(initial=0.0) <= 2e-5
    assert np.all(got[rows:] == 0)
    # identical to the same layer on the materialised concatenation
    cat = torch.from_numpy(np.concatenate([lo, hi], axis=1)).cuda()
    same = lin.rows(cat, torch.tensor([rows], dtype=torch.int32, device="cuda"), activation=act)
    assert torch.equal(same, out)


@pytest.mark.parametrize("rows", [0, 333, 1024])
@pytest.mark.parametrize("with_add", [False, True])
def test_linear_rows_splitk(rows, with_add):
    """FFN second linear 384 -> 192 in split-K form: one launch, the parts sum to x W^T + b (+ residual rows)."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    rng = np.random.default_rng(rows + with_add)
    cap = 1024
    W = (rng.standard_normal((192, 384)) * 0.05).astype(np.float32)
    b = (rng.standard_normal(192) * 0.1).astype(np.float32)
    x = rng.standard_normal((cap, 384)).astype(np.float32)
    add = rng.standard_normal((cap, 192)).astype(np.float32)
    lin = capi.Linear(W, b, precision=capi.DSVT_ATTN_FP32_TC)
    parts = lin.rows_splitk(torch.from_numpy(x).cuda(), torch.tensor([rows], dtype=torch.int32, device="cuda"),
                            add=torch.from_numpy(add).cuda() if with_add else None)
    assert parts.shape == (2, cap, 192)
    got = (parts[0, :rows] + parts[1, :rows]).cpu().numpy()
    ref = x[:rows].astype(np.float64) @ W.T.astype(np.float64) + b + (add[:rows] if with_add else 0.0)
    assert np.abs(got - ref).max(initial=0.0) <= 2e-5


@pytest.mark.parametrize("rows", [1, 127, 128, 1000, 1300])
def test_pos_embed_mlp_one_kernel(rows):
    """dsvt_pos_embed_mlp_launch (Linear(2->192)+BN+ReLU generated inside the Linear(192->192) GEMM) is bit-identical to
    the two launches it replaces, and within the FP32 tolerance of a float64 evaluation (src/dsvt-ai-trt.cpp:461-492)."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    rng = np.random.default_rng(rows)
    cap, C = 1300, 192
    x2 = np.zeros((cap, 2), np.float32)
    x2[:rows] = rng.integers(-12, 12, (rows, 2)).astype(np.float32) + 0.5        # in-window coordinates: x - W/2
    w1 = (rng.standard_normal((C, 2)) * 0.3).astype(np.float32)
    sc = (0.5 + 0.1 * rng.standard_normal(C)).astype(np.float32)
    sh = (0.1 * rng.standard_normal(C)).astype(np.float32)
    w2 = (rng.standard_normal((C, C)) * 0.07).astype(np.float32)
    b2 = (rng.standard_normal(C) * 0.02).astype(np.float32)
    first, second = capi.SmallLinear(w1, sc, sh), capi.Linear(w2, b2, precision=capi.DSVT_ATTN_FP32_TC)
    dx, n = torch.from_numpy(x2).cuda(), torch.tensor([rows], dtype=torch.int32, device="cuda")
    hidden = first(dx, n, activation=2)
    two = second.rows(hidden, n)
    one = torch.full((cap, C), float("nan"), device="cuda")
    capi.pos_embed_mlp(first, second, dx, n, out=one)
    torch.cuda.synchronize()
    assert torch.equal(one, two)
    assert float(one[rows:].abs().max()) == 0.0 if rows < cap else True
    h64 = np.maximum((x2[:rows].astype(np.float64) @ w1.T.astype(np.float64)) * sc + sh, 0.0)
    ref = h64 @ w2.T.astype(np.float64) + b2
    assert np.abs(one[:rows].cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("K,n_ln", [(192, 1), (384, 2), (384, 3)])
@pytest.mark.parametrize("rows", [1, 200, 1000])
def test_linear_rows_norm(K, n_ln, rows):
    """dsvt_linear_rows_norm_launch: linear + a chain of (residual add + LayerNorm) stages in the GEMM epilogue against the
    separate launches (same arithmetic per stage) and a float64 evaluation."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    rng = np.random.default_rng(K + n_ln + rows)
    cap, C = 1300, 192
    x = np.zeros((cap, K), np.float32)
    x[:rows] = rng.standard_normal((rows, K))
    W = (rng.standard_normal((C, K)) * 0.06).astype(np.float32)
    b = (rng.standard_normal(C) * 0.05).astype(np.float32)
    lin = capi.Linear(W, b, precision=capi.DSVT_ATTN_FP32_TC)
    res = [torch.from_numpy(rng.standard_normal((cap, C)).astype(np.float32)).cuda() for _ in range(n_ln)]
    gam = [torch.from_numpy((1 + 0.1 * rng.standard_normal(C)).astype(np.float32)).cuda() for _ in range(n_ln)]
    bet = [torch.from_numpy((0.1 * rng.standard_normal(C)).astype(np.float32)).cuda() for _ in range(n_ln)]
    dx, n = torch.from_numpy(x).cuda(), torch.tensor([rows], dtype=torch.int32, device="cuda")
    stages = list(zip(res, gam, bet))
    out = torch.full((cap, C), float("nan"), device="cuda")
    lin.rows_norm(dx, n, stages, 0.0, out=out)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.all(got[rows:] == 0)
    y = x[:rows].astype(np.float64) @ W.T.astype(np.float64) + b
    for r, g_, b_ in stages:
        y = y + r[:rows].cpu().numpy().astype(np.float64)
        mu = y.mean(1, keepdims=True)
        y = (y - mu) / np.sqrt(((y - mu) ** 2).mean(1, keepdims=True)) * g_.cpu().numpy() + b_.cpu().numpy()
    assert np.abs(got[:rows] - y).max() <= 5e-5
    if K == 192:      # against the separate launches: linear, then the LayerNorm chain kernel (same arithmetic per stage up to
        two = capi.layer_norm_chain(lin.rows(dx, n), n, stages, 0.0)       # the epilogue's reciprocal-multiply: a few ulp)
        assert (out - two).abs().max().item() <= 2e-6


@pytest.mark.parametrize("n_ln", [1, 2, 3])
@pytest.mark.parametrize("rows", [1, 130, 1000, 2999])
def test_ffn_fused(n_ln, rows):
    """dsvt_ffn_fused_launch (FC 192->384, GELU, FC 384->192 and the LayerNorm chain in one kernel, hidden rows in tensor
    memory) against the two-kernel form it replaces (same products in the same order) and a float64 evaluation."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    rng = np.random.default_rng(17 * n_ln + rows)
    cap, C, F = 3000, 192, 384
    x = np.zeros((cap, C), np.float32)
    x[:rows] = rng.standard_normal((rows, C))
    W1 = (rng.standard_normal((F, C)) * 0.07).astype(np.float32)
    b1 = (rng.standard_normal(F) * 0.05).astype(np.float32)
    W2 = (rng.standard_normal((C, F)) * 0.05).astype(np.float32)
    b2 = (rng.standard_normal(C) * 0.05).astype(np.float32)
    fc1 = capi.Linear(W1, b1, precision=capi.DSVT_ATTN_FP32_TC)
    fc2 = capi.Linear(W2, b2, precision=capi.DSVT_ATTN_FP32_TC)
    res = [torch.from_numpy(rng.standard_normal((cap, C)).astype(np.float32)).cuda() for _ in range(n_ln)]
    gam = [torch.from_numpy((1 + 0.1 * rng.standard_normal(C)).astype(np.float32)).cuda() for _ in range(n_ln)]
    bet = [torch.from_numpy((0.1 * rng.standard_normal(C)).astype(np.float32)).cuda() for _ in range(n_ln)]
    dx, n = torch.from_numpy(x).cuda(), torch.tensor([rows], dtype=torch.int32, device="cuda")
    stages = list(zip(res, gam, bet))
    out = torch.full((cap, C), float("nan"), device="cuda")
    fc1.ffn_norm(fc2, dx, n, stages, 0.0, out=out)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.isfinite(got[:rows]).all() and np.all(got[rows:] == 0)
    hidden = fc1.rows(dx, n, activation=1, zero_tails=0)
    two = fc2.rows_norm(hidden, n, stages, 0.0)
    torch.cuda.synchronize()
    assert (out[:rows] - two[:rows]).abs().max().item() <= 2e-6
    h = x[:rows].astype(np.float64) @ W1.T.astype(np.float64) + b1
    h = 0.5 * h * (1.0 + np.tanh(0.7978845608028654 * (h + 0.044715 * h ** 3)))
    y = h @ W2.T.astype(np.float64) + b2
    for r, g_, b_ in stages:
        y = y + r[:rows].cpu().numpy().astype(np.float64)
        mu = y.mean(1, keepdims=True)
        y = (y - mu) / np.sqrt(((y - mu) ** 2).mean(1, keepdims=True)) * g_.cpu().numpy() + b_.cpu().numpy()
    assert np.abs(got[:rows] - y).max() <= 5e-5


def test_pos_embed_mlp_batch_equals_single_launches():
    """dsvt_pos_embed_mlp_batch_launch (up to eight MLPs as the roles of one launch) against one launch per MLP: bit for bit."""
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    rng = np.random.default_rng(5)
    cap, rows, n = 2000, 1500, 8
    firsts, seconds, xs = [], [], []
    for i in range(n):
        firsts.append(capi.SmallLinear((rng.standard_normal((192, 2)) * 0.5).astype(np.float32),
                                       (1 + 0.1 * rng.standard_normal(192)).astype(np.float32), (0.1 * rng.standard_normal(192)).astype(np.float32)))
        seconds.append(capi.Linear((rng.standard_normal((192, 192)) * 0.06).astype(np.float32), (rng.standard_normal(192) * 0.05).astype(np.float32),
                                   precision=capi.DSVT_ATTN_FP32_TC))
        xs.append(torch.from_numpy(rng.integers(0, 24, size=(cap, 2)).astype(np.float32)).cuda())
    nrows = torch.tensor([rows], dtype=torch.int32, device="cuda")
    outs = [torch.full((cap, 192), float("nan"), device="cuda") for _ in range(n)]
    capi.pos_embed_mlp_batch(firsts, seconds, xs, nrows, outs, zero_tails=1)
    torch.cuda.synchronize()
    for i in range(n):
        one = capi.pos_embed_mlp(firsts[i], seconds[i], xs[i], nrows, zero_tails=1)
        assert torch.equal(one, outs[i]) and bool((outs[i][rows:] == 0).all())
