"""Phase timestamps (SM cycles) of the fused QKV + core kernel: CTA 7, its SECOND unit (steady state), bench frame 0.
    make -C dsvt-ai-trt_b200/csrc OUT=$PWD/dsvt-ai-trt_b200/lib_prof EXTRA_DEFS=-DDSVT_PROFILE
    DSVT_B200_LIBDIR=$PWD/dsvt-ai-trt_b200/lib_prof python tools/fused_profile.py"""
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
V, x, gs = f.vox.pillar_num, f.max_voxel[-1], f.gs[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.zero_(); torch.cuda.synchronize()
capi._lib().dsvt_debug_split_profile_reset()
capi.set_attention_fused(w.attn[0], x, f.pos_out[0][0], gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num, V, axis=0,
                         out=f.attn_out, precision=f.precision, workspace=f.attn_ws, plan=f.plans[(0, 0)])
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
capi._lib().dsvt_debug_split_profile(buf)
t = np.array(buf[:], dtype=np.int64)
print(f"CTA 7: {t[30]} units, {t[32]} cycles in total; per-role interval sums (cycles):")
lab = {0: "core: metadata wait + bias loads", 1: "core: waiting for the accumulators", 2: "core: drain Q / K / V", 3: "core: tile barrier",
       4: "core: attention (thread 0)", 5: "core: end barrier", 8: "prod: metadata", 9: "prod: metadata barrier", 10: "prod: load issue",
       11: "prod: waiting for a free stage", 12: "prod: convert + store (incl. load wait)", 16: "issuer: waiting for drained accumulators",
       17: "issuer: waiting for weight chunks", 18: "issuer: waiting for A chunks", 19: "issuer: MMA issue"}
for i in sorted(lab):
    print(f"  {lab[i]:44s} {t[i]:9d}")
