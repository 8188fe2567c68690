"""BASELINE.json configs[4]: density sweep of the voxeliser (a1) and getSet (a2).

For P in {50k..300k} (ring-lidar and the uniform-disc stress) run the voxeliser on a BATCH of frames in one
launch sequence (the regime where an HBM roofline is meaningful: a single 200k-point frame moves only ~11 MB,
below one launch latency) and report algorithmic GB/s = (16 P + 44 Pc + 20 V + 8) x frames / time against the
measured copy bandwidth.  getSet is timed for set in {24, 36, 48}.  Output: one JSON line per case.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=5, flush=None):
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--points", type=int, nargs="*", default=[50000, 100000, 150000, 200000, 250000, 300000])
    ap.add_argument("--tight", action="store_true", help="capacities sized to the data (no zero-tail traffic)")
    args = ap.parse_args()
    pkg = importlib.import_module("dsvt-ai-trt_b200")
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    B = args.batch
    for gen in ("ring_lidar", "uniform_disc"):
        for P in args.points:
            clouds = [getattr(pkg.synth, gen)(P, seed=s) for s in range(min(B, 8))]
            base = pkg.config.WAYMO.with_(max_pillars_num=110000, max_win_num=8192)
            cfg = base.with_(max_points_num=P) if args.tight else base
            pts = np.zeros((B, cfg.max_points_num, 4), np.float32)
            for i in range(B):
                pts[i, :P] = clouds[i % len(clouds)]
            d_pts = torch.from_numpy(pts).cuda()
            sizes = torch.full((B,), P, dtype=torch.int32, device="cuda")
            for zt in (1, 0):
                vox = capi.Points2Features(cfg, batch=B, zero_tails=zt)
                vox(d_pts, sizes); torch.cuda.synchronize()
                us = timed(lambda: vox(d_pts, sizes), flush=flush)
                V = vox.pillar_num.cpu().numpy().astype(np.int64)
                Pc = vox.point_num.cpu().numpy().astype(np.int64)
                alg = int((16 * P + 44 * Pc + 20 * V + 8).sum())
                contract = int(B * (16 * P + 40 * cfg.max_points_num_voxel_filter + 212 * cfg.max_pillars_num + 8))
                gbs = alg / us * 1e-3
                print(json.dumps({"kernel": "points2features", "cloud": gen, "points": P, "frames": B, "zero_tails": zt,
                                  "pillars": int(V[0]), "kept": int(Pc[0]), "us": round(us, 1),
                                  "us_per_frame": round(us / B, 2), "algorithmic_bytes": alg, "contract_bytes": contract,
                                  "algorithmic_gbs": round(gbs, 1), "frac_of_measured_copy_bw": round(gbs / peaks["hbm_gbs"], 4)}))
                del vox
            # getSet for set sizes 24 / 36 / 48 on frame 0's partition
            v1 = capi.Points2Features(cfg, batch=1)
            v1(d_pts[:1], sizes[:1])
            for S in (24, 36, 48):
                c2 = cfg.with_(voxel_num_set=S)
                wp = capi.WindowPartition(c2, 0)
                wp(v1.coords, v1.pillar_num)
                gs = capi.GetSet(c2, 0)
                gs(wp.global_index, wp.coors_in_win, wp.voxel_num_in_win, wp.win_num); torch.cuda.synchronize()
                us = timed(lambda: gs(wp.global_index, wp.coors_in_win, wp.voxel_num_in_win, wp.win_num), flush=flush)
                Vn, W, Ns = int(v1.pillar_num[0]), int(wp.win_num[0]), int(gs.set_num[0])
                alg = 16 * Vn + 4 * W + S * 4 * 20 * Ns + 4
                print(json.dumps({"kernel": "get_set", "cloud": gen, "points": P, "set": S, "pillars": Vn, "windows": W,
                                  "sets": Ns, "us": round(us, 1), "algorithmic_bytes": alg,
                                  "algorithmic_gbs": round(alg / us * 1e-3, 1)}))


if __name__ == "__main__":
    main()
