"""Timing of the fused FFN kernel against the two kernels it replaces, on the bench's frame 0 (cold L2, median of 9)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
cfg = pkg.config.WAYMO
w = pipeline.FrameWeights(cfg, seed=0)
f = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC, ffn="epilogue", backbone=True)
f.load_points(pkg.synth.ring_lidar(200000, seed=0))
f.run(); torch.cuda.synchronize()
V = f.vox.pillar_num
fc1, fc2 = w.ffn[0]
x = f.blk_out[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush_r = torch.zeros(64 << 20, dtype=torch.int32, device="cuda")
def timed(fn, reps=9):
    ts = []
    for _ in range(reps):
        flush.zero_(); flush_r.max()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]
for n_ln in (2, 3):
    st = [(f.src, w.gamma[1], w.beta[1]), (x, w.gamma[2], w.beta[2]), (x, w.gamma[3], w.beta[3])][:n_ln]
    o1, o2 = torch.empty_like(f.src), torch.empty_like(f.src)
    def two():
        fc1.rows(f.src, V, activation=1, out=f.gelu_out, zero_tails=0)
        fc2.rows_norm(f.gelu_out, V, st, cfg.layer_norm_eps, out=o1)
    one = lambda: fc1.ffn_norm(fc2, f.src, V, st, cfg.layer_norm_eps, out=o2)
    two(); one(); torch.cuda.synchronize()
    n = int(V[0])
    print(f"n_ln={n_ln}: max |fused - two kernels| = {(o1[:n] - o2[:n]).abs().max().item():.3g}, two kernels {timed(two):.1f} us, fused {timed(one):.1f} us")

if os.environ.get("DSVT_B200_LIBDIR", "").endswith("lib_prof"):       # phase stamps of one CTA (tile 20), SM cycles
    import ctypes, numpy as np
    st = [(f.src, w.gamma[1], w.beta[1]), (x, w.gamma[2], w.beta[2])]
    flush.zero_(); flush_r.max(); torch.cuda.synchronize()
    capi._lib().dsvt_debug_split_profile_reset() if hasattr(capi._lib(), "dsvt_debug_split_profile_reset") else None
    if "--tail" in sys.argv:      # the attention-tail form: out-projection + norm1 in front (needs the core's rows in the workspace)
        gs = f.gs[0]
        capi.set_attention_fused(w.attn[0], x, f.pos_out[0][0], gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num, V, axis=0,
                                 out=f.src, precision=f.precision, workspace=f.attn_ws, plan=f.plans[(0, 0)], stages=3)
        torch.cuda.synchronize(); flush.zero_(); flush_r.max(); torch.cuda.synchronize()
        capi.attention_tail_ffn(w.attn[0], fc1, fc2, x, gs.global_index_in_set[0], V, 0, f.plans[(0, 0)], f.attn_ws,
                                (w.gamma[0], w.beta[0], cfg.layer_norm_eps), st, cfg.layer_norm_eps, src=f.src, out=o2)
    else:
        fc1.ffn_norm(fc2, f.src, V, st, cfg.layer_norm_eps, out=o2)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)(); capi._lib().dsvt_debug_split_profile(buf)
    t = np.array(buf[:], dtype=np.int64)
    lab = {0: "start", 1: "setup done", 2: "x (or o) image staged", 26: "[tail form] G_o complete", 27: "[tail form] norm1 done, src image staged", 9: "ACC2 complete", 10: "LN pass A done", 11: "LN pass B half", 13: "CTA end"}
    for p in range(6): lab[3 + p] = f"workers: A2({p}) written"
    ops = ["G1(0)", "G1(1)", "G2(0)", "G1(2)", "G2(1)", "G1(3)", "G2(2)", "G1(4)", "G2(3)", "G1(5)", "G2(4)", "G2(5)"]
    for i, o in enumerate(ops): lab[14 + i] = f"issuer: begins {o}"
    for i in sorted(lab, key=lambda i: t[i]): print(f"  {lab[i]:30s} t={t[i] - t[0]:7d}")
