"""Times one set-attention call (bench frame 0, both partitions) for every precision mode and prints the max-abs
difference against the CUDA-core FP32 kernel.  CUDA events, L2 flushed between calls."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("dsvt-ai-trt_b200"); capi = importlib.import_module("dsvt-ai-trt_b200.capi")
pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
cfg = pkg.config.WAYMO
points = 200000
noflush = "--noflush" in sys.argv
w = pipeline.FrameWeights(cfg)
f = pipeline.HotPathFrame(cfg, w)
f.load_points(pkg.synth.ring_lidar(points, 0))
f.run(); torch.cuda.synchronize()
V = int(f.vox.pillar_num[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = {0: "fp32 cuda-core", 2: "fp16 fused tcgen05", 3: "fp32-accurate split GEMM", 4: "fp16 GEMM pipeline"}
ws = torch.empty(capi.set_attention_workspace_bytes(1, cfg.max_win_num, 36, 192, 8, cfg.max_pillars_num, 3), dtype=torch.uint8, device="cuda")
for which in (0, 1):
    gs = f.gs[which]
    ns = int(gs.set_num[0])
    ref = None
    plan = capi.set_attention_plan(gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num, 0, cfg.max_pillars_num)
    for prec, planned in ((0, False), (2, False), (3, False), (3, True), (4, True)):
        out = torch.zeros_like(f.attn_out)
        call = lambda: capi.set_attention_fused(w.attn[0], f.x0, f.pos[0][0], gs.global_index_in_set[0], gs.mask_expand_0[0],
                                                gs.set_num, f.vox.pillar_num, axis=0, out=out, precision=prec, workspace=ws,
                                                plan=plan if planned else None)
        for _ in range(3): call()
        ts = []
        for _ in range(20):
            if not noflush: flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        if prec == 0: ref = out.clone()
        err = (out[:V] - ref[:V]).abs().max().item()
        stage = ""
        if prec in (3, 4):
            import ctypes
            lib = capi._lib(); lib.dsvt_debug_attention_stage_timing(1)
            flush.zero_(); call(); torch.cuda.synchronize()
            buf = (ctypes.c_float * 3)(); lib.dsvt_debug_attention_stage_us(buf); lib.dsvt_debug_attention_stage_timing(0)
            stage = f"  stages [plan+]qkv/core/out = {buf[0]:.1f}/{buf[1]:.1f}/{buf[2]:.1f} us" + (" (plan reused)" if planned else " (plan built per call)")
        print(f"partition {which}: {V} voxels {ns} sets  {names[prec]:26s} median {np.median(ts):8.1f} us  min {min(ts):8.1f} us  max|d| vs fp32 {err:.3e}{stage}")
