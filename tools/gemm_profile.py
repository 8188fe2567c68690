"""One eager pass over the GEMM-class launches of the headline frame (bench frame 0) between cudaProfilerStart/Stop,
for `ncu --profile-from-start off` captures:

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'proj_tile|attn_core' \
        -o gpurun_out/r2_gemm python tools/gemm_profile.py
"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

pkg = importlib.import_module("dsvt-ai-trt_b200")
capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
cfg = pkg.config.WAYMO
w = pipeline.FrameWeights(cfg, seed=0)
f = pipeline.HotPathFrame(cfg, w, precision=capi.DSVT_ATTN_FP32_TC, ffn="epilogue", backbone=True)
f.load_points(pkg.synth.ring_lidar(200000, seed=0))
for _ in range(2):
    f.run()
torch.cuda.synchronize()
V = f.vox.pillar_num
x = f.max_voxel[-1]
gs = f.gs[0]
fc1, fc2 = w.ffn[0]
first, second = w.glue["pos"][0][0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush.zero_()
torch.cuda.synchronize()
torch.cuda.profiler.start()
capi.set_attention_fused(w.attn[0], x, f.pos_out[0][0], gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num, V, axis=0,
                         out=f.src, precision=f.precision, workspace=f.attn_ws, plan=f.plans[(0, 0)],
                         norm=(x, w.gamma[0], w.beta[0], cfg.layer_norm_eps))
fc1.rows(f.src, V, activation=1, out=f.gelu_out, zero_tails=0)
fc2.rows_norm(f.gelu_out, V, [(f.src, w.gamma[1], w.beta[1]), (x, w.gamma[2], w.beta[2])], cfg.layer_norm_eps, out=f.x_a)
capi.pos_embed_mlp(first, second, f.wp[0].coors_in_win_x_y[0], V, out=f.pos_out[0][0], zero_tails=0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled: QKV GEMM, core, out-projection + norm1, FFN linear 1 (+GELU), FFN linear 2 + LayerNorm chain, pos-embed MLP")
