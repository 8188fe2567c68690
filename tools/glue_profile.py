"""Times TorchScatterMax (F = 96, 192) and Map2Bev on the bench frame (CUDA events, L2 flushed)."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
cfg = pkg.config.WAYMO
w = pipeline.FrameWeights(cfg); f = pipeline.HotPathFrame(cfg, w, precision=3)
f.load_points(pkg.synth.ring_lidar(200000, 0)); f.run(); torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
V, Pc = int(f.vox.pillar_num[0]), int(f.vox.point_num[0])
def timed(fn):
    ts = []
    for _ in range(7):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[3]
for k, F in enumerate(cfg.pfn_channels):
    for pn in (f.vox.point_num, None):
        us = timed(lambda: capi.torch_scatter_max(w.pfn_out[k], f.vox.point_index_in_voxel[0], f.vox.point_num_in_voxel[0], f.vox.pillar_num, pn, max_point=f.max_point[k], max_voxel=f.max_voxel[k]))
        alg = 4 * F * (2 * Pc + V); con = 4 * F * (Pc + cfg.max_points_num_voxel_filter + cfg.max_pillars_num)
        print(f"scatter_max F={F} point_num={'given' if pn is not None else 'absent (full clear)'}: {us:.1f} us  algorithmic {alg/us*1e-3:.0f} GB/s  contract {con/us*1e-3:.0f} GB/s")
us = timed(lambda: capi.map2bev(f.final, f.vox.coords[0], f.vox.pillar_num, cfg.grid_x, cfg.grid_y, out=f.bev))
print(f"map2bev: {us:.1f} us  contract {(4*192*(cfg.grid_x*cfg.grid_y+V))/us*1e-3:.0f} GB/s")
