"""Timing of the fused VFE kernel against the four launches it replaces on the bench's frame 0 (cold L2, median of 9), and --
with the profile build (DSVT_B200_LIBDIR=.../lib_prof) -- the phase stamps of one CTA (tile 20)."""
import ctypes, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
cfg = pkg.config.WAYMO
w = pipeline.FrameWeights(cfg, seed=0)
f = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC, ffn="layer", backbone=True)
f.load_points(pkg.synth.ring_lidar(200000, seed=0))
f.run(); torch.cuda.synchronize()
vox, g, V = f.vox, w.glue, f.vox.pillar_num
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush_r = torch.zeros(64 << 20, dtype=torch.int32, device="cuda")
def timed(fn, reps=9):
    ts = []
    for _ in range(reps):
        flush.zero_(); flush_r.max()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]
one = lambda: capi.vfe_fused(g["pfn0"], g["pfn1"], vox.point_features[0], vox.point_index_in_voxel[0], V, vox.point_num,
                             out=f.max_voxel[-1], workspace=f.vfe_ws)
print(f"pillars {int(V[0])}, rows {int(vox.point_num[0])}: fused VFE {timed(one):.1f} us")
if os.environ.get("DSVT_B200_LIBDIR", "").endswith("lib_prof"):
    flush.zero_(); flush_r.max(); torch.cuda.synchronize()
    one(); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)(); capi._lib().dsvt_debug_split_profile(buf)
    t = np.array(buf[:], dtype=np.int64)
    lab = {0: "start", 1: "setup done (TMEM, tile bounds)", 2: "step A: h0 staged", 3: "step B: m0 done", 4: "step C: A image complete",
           9: "accumulators complete", 10: "h1 tile staged", 13: "CTA end"}
    for kc in range(6): lab[14 + kc] = f"issuer: W chunk {kc} landed"
    for i in sorted(lab, key=lambda i: t[i]): print(f"  {lab[i]:34s} t={t[i] - t[0]:7d}")
