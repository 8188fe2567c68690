"""One bench frame (Waymo capacities, 200 k-point cloud, FFN linears executed) for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/memcheck_frame.py [precision] [--backbone]
(also the command of the committed ncu launch list of the complete 3-D backbone frame)"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
backbone = "--backbone" in sys.argv          # every layer of the 3-D backbone (PFN, position embedding, FFN) as one data flow
prec = int(args[0]) if args else capi.DSVT_ATTN_FP32_TC
cfg = pkg.config.WAYMO
w = pipeline.FrameWeights(cfg)
for ffn, zt in (("graph", 1), ("kernel", 1), ("layer", 1)):
    f = pipeline.HotPathFrame(cfg, w, precision=prec, ffn=ffn, zero_tails=zt, backbone=backbone)
    f.load_points(pkg.synth.ring_lidar(200000, 0))
    f.run(); torch.cuda.synchronize()
    print(f"ffn={ffn} zero_tails={zt}: pillars {int(f.vox.pillar_num[0])}, sets {int(f.gs[0].set_num[0])}/{int(f.gs[1].set_num[0])}, "
          f"boxes {int(f.valid[0])}, launches {f.launches_per_frame}, finite {bool(torch.isfinite(f.final[: int(f.vox.pillar_num[0])]).all())}")
