"""Phase stamps (SM cycles) of one CTA (tile 20) of the tile-wide QKV projection kernel on the bench's frame 0 and its isolated
time.  Stamps need the profile build (DSVT_B200_LIBDIR=.../lib_prof)."""
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
V, gs, x = f.vox.pillar_num, f.gs[0], f.blk_out[0]
attn = lambda stages: capi.set_attention_fused(
    w.attn[0], x, f.attn_pos(0, 0)[0], gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num, V, axis=0, out=f.src_b,
    precision=f.precision, workspace=f.attn_ws, plan=f.plans[(0, 0)], norm=(x, w.gamma[0], w.beta[0], cfg.layer_norm_eps), stages=stages,
    pos_table=f.attn_pos(0, 0)[1])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush_r = torch.zeros(64 << 20, dtype=torch.int32, device="cuda")
ts = []
for _ in range(9):
    flush.zero_(); flush_r.max()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); attn(1); b.record(); b.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
print(f"QKV projection, isolated (cold L2): {sorted(ts)[4]:.1f} us")
if os.environ.get("DSVT_B200_LIBDIR", "").endswith("lib_prof"):
    flush.zero_(); flush_r.max(); torch.cuda.synchronize()
    attn(1); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)(); capi._lib().dsvt_debug_split_profile(buf)
    t = np.array(buf[:], dtype=np.int64)
    lab = {0: "start", 1: "setup done", 2: "x + pos image staged", 3: "Q drained", 4: "x image staged", 5: "K drained", 6: "V drained", 13: "CTA end"}
    for kc in range(6): lab[14 + kc] = f"issuer: G_q chunk {kc} (A + W landed)"
    for i in sorted(lab, key=lambda i: t[i]): print(f"  {lab[i]:36s} t={t[i] - t[0]:7d}")
