"""Phase timestamps (SM cycles) of one CTA (set 5) of the FP32 set-attention kernel on the bench's frame 0."""
import ctypes, importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
cfg = pkg.config.WAYMO
w = pipeline.FrameWeights(cfg)
f = pipeline.HotPathFrame(cfg, w)
f.load_points(pkg.synth.ring_lidar(200000, 0))
f.run(); torch.cuda.synchronize()
gs = f.gs[0]
for _ in range(3):
    capi.set_attention_fused(w.attn[0], f.x0, f.pos[0][0], gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num, f.vox.pillar_num, axis=0, out=f.attn_out, precision=0)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
capi._lib().dsvt_debug_fp32_profile(buf)
t = np.array(buf[:], dtype=np.int64)
names = ["start", "compaction", "tile qk loaded", "proj Q", "proj K", "tile v loaded", "proj V", "scores g0", "softmax g0", "PV g0", "attention done (both groups)", "out-proj + scatter"]
print("unique tokens in this set:", t[20])
for i in range(1, 12):
    print(f"{names[i]:32s} +{t[i]-t[i-1]:7d}  (t={t[i]-t[0]})")
