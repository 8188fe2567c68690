"""Phase timestamps (SM cycles) of ONE CTA (tile 20, role 0) of the single-pass tile GEMM on the bench's frame 0, for each
of its uses.  Needs the profile build:
    make -C dsvt-ai-trt_b200/csrc OUT=$PWD/dsvt-ai-trt_b200/lib_prof EXTRA_DEFS=-DDSVT_PROFILE
    DSVT_B200_LIBDIR=$PWD/dsvt-ai-trt_b200/lib_prof python tools/tile_profile.py
"""
import ctypes, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
cfg = pkg.config.WAYMO
w = pipeline.FrameWeights(cfg, seed=0)
f = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC, ffn="epilogue", backbone=True)
f.load_points(pkg.synth.ring_lidar(200000, seed=0))
f.run(); f.run(); torch.cuda.synchronize()
V = f.vox.pillar_num
fc1, fc2 = w.ffn[0]
first, second = w.glue["pos"][0][0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush_r = torch.zeros(64 << 20, dtype=torch.int32, device="cuda")
lab = {0: "start", 1: "setup done (barriers, TMEM)", 8: "producer: all chunks staged", 9: "accumulators complete", 10: "slab 0 stored",
       11: "slab 1 stored", 12: "slab 2 stored", 13: "CTA end"}
for kc in range(6):
    lab[2 + kc] = f"producer: chunk {kc} staged"; lab[14 + kc] = f"issuer: W chunk {kc} landed"; lab[20 + kc] = f"issuer: A chunk {kc} full"
def show(title, fn):
    flush.zero_(); flush_r.max(); torch.cuda.synchronize()
    fn(); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)()
    capi._lib().dsvt_debug_split_profile(buf)
    t = np.array(buf[:], dtype=np.int64)
    print(title)
    for i in sorted(lab, key=lambda i: t[i]):
        print(f"  {lab[i]:34s} t={t[i]-t[0]:7d}")
x = f.blk_out[0]
gs = f.gs[0]
plan = f.plans[(0, 0)]
ln = [(x, w.gamma[1], w.beta[1])]
attn = lambda stages: capi.set_attention_fused(
    w.attn[0], x, f.pos_out[0][0], gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num, V, axis=0, out=f.src_b,
    precision=f.precision, workspace=f.attn_ws, plan=plan, norm=(x, w.gamma[0], w.beta[0], cfg.layer_norm_eps), stages=stages)
show("plain linear 192->192 (the pos-embed MLP's second layer, A read from memory; cold L2):", lambda: second.rows(f.pos_hidden, V, out=f.pos_out[0][0], zero_tails=0))
show("QKV projection GEMM, role q (A = x + pos):", lambda: attn(1))
attn(2); torch.cuda.synchronize()
show("out-projection + norm1 (LayerNorm epilogue):", lambda: attn(4))
show("FFN linear 1 + GELU (192->384):", lambda: fc1.rows(f.src, V, activation=1, out=f.gelu_out, zero_tails=0))
show("FFN linear 2 (K = 384) + norm2 (LayerNorm epilogue):", lambda: fc2.rows_norm(f.gelu_out, V, ln, cfg.layer_norm_eps, out=f.src_b))
