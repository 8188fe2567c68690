"""Phase timestamps (SM cycles) of one CTA of the GEMM-pipeline attention kernels on the bench's frame 0."""
import ctypes, importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
cfg = pkg.config.WAYMO
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
w = pipeline.FrameWeights(cfg)
f = pipeline.HotPathFrame(cfg, w, precision=prec)
f.load_points(pkg.synth.ring_lidar(200000, 0))
f.run(); torch.cuda.synchronize()
gs = f.gs[0]
for _ in range(3):
    capi.set_attention_fused(w.attn[0], f.x0, f.pos[0][0], gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num,
                             f.vox.pillar_num, axis=0, out=f.attn_out, precision=prec, workspace=f.attn_ws)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
capi._lib().dsvt_debug_split_profile(buf)
t = np.array(buf[:], dtype=np.int64)
def show(base, title):
    lab = {0: "start", 1: "setup done (barriers, TMEM)", 14: "producers: all steps staged", 15: "issuer: all MMAs issued",
           16: "epilogue: tile 0 stored", 17: "epilogue: last valid tile stored", 20: "epilogue: tile 0 accumulators ready", 21: "CTA end"}
    for sl in range(3):          # -DDSVT_EPI_PROBE builds only (QKV kernel)
        if base == 0 and t[22 + sl * 3]:
            lab[22 + sl * 3] = f"epilogue: slab {sl} tmem loaded"; lab[23 + sl * 3] = f"epilogue: slab {sl} transposed"
            lab[24 + sl * 3] = f"epilogue: slab {sl} stored"
    for kc in range(6):
        lab[2 + kc] = f"producer: chunk {kc} staged"; lab[8 + kc] = f"issuer: chunk {kc} full (tile 0)"
    print(title)
    for i in sorted(lab, key=lambda i: t[base + i]):
        print(f"  {lab[i]:38s} t={t[base + i]-t[base]:7d}")
show(0, "QKV projection GEMM, CTA 3 (role Q, tiles 1, 50, 99):")
show(40, "out-projection GEMM, CTA 3 (tile 3 + tail tiles):")
c = t[32:36]
print("core kernel, set 5:", "compaction", c[1]-c[0], "K/V staged", c[2]-c[0], "done", c[3]-c[0])

if t[38] > 0 and t[38] < 10**7:
    print(f"core kernel, sums over {t[38]} set iterations (all launches so far): avg staging {t[36]/t[38]:.0f} cycles, avg compute+store {t[37]/t[38]:.0f} cycles, avg tokens {t[39]/t[38]:.1f}")
