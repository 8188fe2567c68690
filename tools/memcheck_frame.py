"""One bench frame (Waymo capacities, 200 k-point cloud, FFN linears executed) for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/memcheck_frame.py [precision]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
prec = int(sys.argv[1]) if len(sys.argv) > 1 else capi.DSVT_ATTN_FP32_TC
cfg = pkg.config.WAYMO
w = pipeline.FrameWeights(cfg)
for ffn, zt in (("graph", 1), ("fused", 0)):
    f = pipeline.HotPathFrame(cfg, w, precision=prec, ffn=ffn, zero_tails=zt)
    f.load_points(pkg.synth.ring_lidar(200000, 0))
    f.run(); torch.cuda.synchronize()
    print(f"ffn={ffn} zero_tails={zt}: pillars {int(f.vox.pillar_num[0])}, sets {int(f.gs[0].set_num[0])}/{int(f.gs[1].set_num[0])}, "
          f"boxes {int(f.valid[0])}, launches {f.launches_per_frame}, finite {bool(torch.isfinite(f.final[: int(f.vox.pillar_num[0])]).all())}")
