"""One HEADLINE frame (bench.py's frame kind "backbone3d": Waymo capacities, 200 k-point ring cloud seed 0, every layer of the
3-D backbone, norms / GELU in the GEMM epilogues, each FFN one kernel) between cudaProfilerStart/Stop, after one warm frame:
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r2_frame python tools/profile_frame.py
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_frame.py
[--head] adds the post-process graph + rotated NMS (frame kind "backbone3d_postprocess")."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
cfg = pkg.config.WAYMO
w = pipeline.FrameWeights(cfg)
f = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC, ffn="layer", backbone=True, head="--head" in sys.argv)
f.load_points(pkg.synth.ring_lidar(200000, 0))
f.run(); torch.cuda.synchronize()
torch.cuda.profiler.start()
f.run(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"pillars {int(f.vox.pillar_num[0])}, sets {int(f.gs[0].set_num[0])}/{int(f.gs[1].set_num[0])}, boxes {int(f.valid[0])}, "
      f"launches {f.launches_per_frame}")
