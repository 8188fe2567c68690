"""Prints the phase timestamps (SM cycles) of the first tile of CTA 0 of the FP16 tensor-core attention kernel."""
import ctypes, importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
capi = importlib.import_module("dsvt-ai-trt_b200.capi")
rng = np.random.default_rng(0)
n_sets, S, C, H = 440, 36, 192, 8
V = n_sets * 13
x = torch.randn(V + 8, C, device="cuda"); pos = torch.randn(V + 8, C, device="cuda")
idx = torch.from_numpy(rng.integers(0, V, (2, n_sets + 4, S)).astype(np.int32)).cuda()
mask = torch.zeros(n_sets + 4, H, S, device="cuda")
W = capi.AttentionWeights(rng.standard_normal((3*C, C)).astype(np.float32)*0.06, np.zeros(3*C, np.float32),
                          rng.standard_normal((C, C)).astype(np.float32)*0.06, np.zeros(C, np.float32))
ns = torch.tensor([n_sets], dtype=torch.int32, device="cuda"); vn = torch.tensor([V], dtype=torch.int32, device="cuda")
for _ in range(3):
    capi.set_attention_fused(W, x, pos, idx, mask, ns, vn, 0, precision=2)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
capi._lib().dsvt_debug_tc_profile(buf)
t = np.array(buf[:], dtype=np.int64)
names = {0: "tile start", 1: "staged", 20: "heads done (PV7)", 21: "O tile written", 22: "out-proj done", 23: "tile end"}
for h in range(2):
    names.update({2+5*h: f"h{h} proj done", 3+5*h: f"h{h} epilogue+sync", 4+5*h: f"h{h} S done", 5+5*h: f"h{h} softmax+sync", 6+5*h: f"h{h} PV/proj issued"})
names.update({34: "  h0 PV issue start", 33: "  h0 PV issued", 30: "  h1 proj: before weight wait", 31: "  h1 proj: weights landed", 32: "  h1 proj issued"})
order = [0,1,2,3,4,5,34,33,30,31,32,6,7,8,9,10,11,20,21,22,23]
prev = t[0]
for i in order:
    print(f"{names[i]:24s} +{t[i]-prev:7d}  (t={t[i]-t[0]})"); prev = t[i]

print("---- warp-specialised kernel (attention_tc2.cu) ----")
capi._lib().dsvt_debug_tc2_profile(buf)
t = np.array(buf[:], dtype=np.int64)
t0 = t[0]
lab = {0: "issuer: tile start", 1: "issuer: A tiles staged", 10: "issuer: PV(7) issued", 11: "issuer: Wout+O tile ready", 12: "issuer: out-proj issued",
       16: "E: staging done", 26: "workers: out-proj done", 27: "workers: tile end"}
for h in range(8):
    lab[2 + h] = f"issuer: step {h} done"; lab[17 + h] = f"E: qkv({h}) written"; lab[32 + h] = f"S: P({h}) written"
for i in sorted(lab, key=lambda i: t[i]):
    print(f"{lab[i]:28s} t={t[i]-t0:7d}")
