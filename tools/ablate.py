"""Marginal cost of each plugin group WITH SEVERAL FRAMES IN FLIGHT (the bench's regime): frames/s of the bench step
with one group left out of every frame's captured graph.  Isolated per-plugin times (bench.py `plugins`) overstate the
latency-bound kernels, which overlap with other frames' work; this shows what each group really costs the step.

    python tools/ablate.py [--steps 5] > gpurun_out/ablate.jsonl
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--groups", default="vox,pfn,smax,part,plan,pos,attn,attn_qkv,attn_core,attn_out,ln1,ffn1,ffn2,lnc,m2b,fbox")
    ap.add_argument("--kind", default="backbone3d", help="bench.FRAME_KINDS key")
    args = ap.parse_args()
    import torch
    bench = importlib.import_module("bench")
    pkg = importlib.import_module("dsvt-ai-trt_b200")
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    cfg = pkg.config.WAYMO
    weights = pipeline.FrameWeights(cfg, seed=0)
    streams = [torch.cuda.Stream() for _ in range(args.streams)]
    slots = []
    for i in range(args.frames):
        s = bench.Slot(pipeline, cfg, weights, capi.DSVT_ATTN_FP32_TC, pkg.synth.ring_lidar(bench.N_POINTS, seed=i), i, kind=args.kind)
        s.frame.run()              # every buffer holds a valid frame before groups are left out
        slots.append(s)
    torch.cuda.synchronize()

    def measure(skip):
        for i, s in enumerate(slots):
            s.frame.skip = frozenset(skip)
            s.capture(streams[i % len(streams)])
        torch.cuda.synchronize()
        bench.run_steps(slots, streams, 3, host=False)
        ms = bench.run_steps(slots, streams, args.steps, host=False)
        return ms / (args.steps * args.frames) * 1e3, slots[0].frame.launches_per_frame   # us per frame

    full, n_full = measure(())
    print(json.dumps({"skip": None, "us_per_frame": round(full, 1), "frames_per_s": round(1e6 / full, 1), "launches": n_full}))
    for g in args.groups.split(","):
        us, n = measure(g.split("+"))
        print(json.dumps({"skip": g, "us_per_frame": round(us, 1), "marginal_us": round(full - us, 1),
                          "share": round((full - us) / full, 3), "launches": n}), flush=True)


if __name__ == "__main__":
    main()
