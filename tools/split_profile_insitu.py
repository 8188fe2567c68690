"""Phase timestamps of the GEMM-pipeline attention kernels captured IN SITU: 8 frames on 4 streams replay their CUDA
graphs (the bench's protocol), then the device-side profile of the last launches is read back."""
import ctypes, importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
import bench
cfg = pkg.config.WAYMO
w = pipeline.FrameWeights(cfg)
streams = [torch.cuda.Stream() for _ in range(4)]
slots = []
for i in range(8):
    s = bench.Slot(pipeline, cfg, w, capi.DSVT_ATTN_FP32_TC, pkg.synth.ring_lidar(200000, i), i)
    s.capture(streams[i % 4]); slots.append(s)
ms = bench.run_steps(slots, streams, 3, host=False)
ms = bench.run_steps(slots, streams, 5, host=False)
print(f"{8 * 5 / (ms * 1e-3):.1f} frames/s")
buf = (ctypes.c_longlong * 64)()
capi._lib().dsvt_debug_split_profile(buf)
t = np.array(buf[:], dtype=np.int64)
def show(base, title):
    lab = {0: "start", 1: "setup done (barriers, TMEM)", 14: "producers: all steps staged", 15: "issuer: all MMAs issued",
           16: "epilogue: tile 0 stored", 17: "epilogue: last valid tile stored", 20: "epilogue: tile 0 accumulators ready", 21: "CTA end"}
    for kc in range(6):
        lab[2 + kc] = f"producer: chunk {kc} staged"; lab[8 + kc] = f"issuer: chunk {kc} full (tile 0)"
    print(title)
    for i in sorted(lab, key=lambda i: t[base + i]):
        print(f"  {lab[i]:38s} t={t[base + i]-t[base]:7d}")
show(0, "QKV projection GEMM, CTA 3:")
show(40, "out-projection GEMM, CTA 3:")
c = t[32:36]
print("core kernel, CTA 5:", "compaction", c[1]-c[0], "K/V staged", c[2]-c[0], "done", c[3]-c[0])
