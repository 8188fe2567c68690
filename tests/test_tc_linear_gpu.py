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
