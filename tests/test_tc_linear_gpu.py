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
