import importlib, sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
cfg = pkg.config.WAYMO
B, P = 16, 200000
pts = np.zeros((B, cfg.max_points_num, 4), np.float32)
for i in range(B): pts[i, :P] = pkg.synth.ring_lidar(P, seed=i % 4)
d = torch.from_numpy(pts).cuda(); sizes = torch.full((B,), P, dtype=torch.int32, device="cuda")
vox = capi.Points2Features(cfg, batch=B, zero_tails=1)
for _ in range(3): vox(d, sizes)
torch.cuda.synchronize()
