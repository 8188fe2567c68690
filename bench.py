#!/usr/bin/env python
"""bench.py -- Waymo-shape frames/sec of the DSVT hot path on N B200s (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of F synthetic 200k-point clouds per GPU
(BASELINE.json configs[1]; frames are independent, one frame per CUDA stream, frame f -> rank f mod N).
The frame is SURVEY.md 8(d)'s "3-D backbone frame": raw points -> voxeliser -> PFN -> scatter-max -> window partition
-> getSet -> 4 DSVT blocks (8 set attentions, 8 FFNs, 28 LayerNorms, 8 position-embedding MLPs) -> BEV map, +
filterBoxByScore on synthetic head candidates -- every layer of the reference's 3-D backbone as ONE data flow.
  value : whole-job frames/s, clouds already resident in HBM, each frame one captured CUDA graph replay.
  e2e   : the same frames through HOST buffers: pinned host cloud -> H2D -> graph -> D2H of the [500,9]
          boxes + count, inside the timed region.
  legs  : the same clouds as (a) the plugin-only frame (round 1's headline: TensorRT-native layers stood in by fixed
          tensors), (b) the reference graph's node structure (FC -> GeluPlugin -> FC), (c) the FP16 configuration
          (BASELINE.json configs[2]), (d) without the contract's zero tails, (e) BASELINE.json configs[3] run literally:
          64 frames x 180k points, frame f -> rank f mod N, one frame per stream, NCCL gather inside the timed region.
  roofline / plugins : per-plugin device time measured with CUDA events in an instrumented pass over the
          same frames, against MEASURED_PEAKS.json.
  cpu_baseline : the CPU oracle port (oracle/dsvt_oracle.c, 1 core) on a bounded sample of one frame.
`--impl reference` times the reference's OWN kernels (its plugin sources compiled unmodified into oracle/_ref,
capacities raised through a params.h placed first on the include path) for the same plugin sequence; the set
attention, which lives inside closed-source TensorRT in the reference, is stood in by PyTorch eager.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "waymo_shape_hot_path_frames_per_sec"
UNIT = "frames/s"
N_POINTS = 200000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=16, help="frames per GPU per step")
    ap.add_argument("--streams", type=int, default=8, help="concurrent frame streams per GPU")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp32_cuda", "fp32_tc", "tf32", "fp16", "fp16_gemm"])
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="headline frame only (skip every extra leg)")
    ap.add_argument("--no-config4", action="store_true", help="skip the BASELINE.json configs[3] leg (64 x 180k frames)")
    ap.add_argument("--no-breakdown", action="store_true", help="skip the per-plugin breakdown of the reference arm")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        # under load = the upper half of the samples (idle samples before/after the region are dropped)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "sm_max_mhz": p.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ---------------------------------------------------------------------------------------------
def dist_setup(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    sharding = importlib.import_module("dsvt-ai-trt_b200.sharding")
    return sharding.reduce_max(x, device="cuda")


# ---------------------------------------------------------------------------------------------
# frame kinds (HotPathFrame keyword arguments); "backbone3d" is the headline
FRAME_KINDS = {
    "backbone3d": dict(ffn="layer", backbone=True),          # every layer of the 3-D backbone; 3 kernels per encoder layer: QKV
                                                             # projection, per-set core, then out-projection + norm1 + FFN (hidden
                                                             # rows in tensor memory) + the LayerNorm chain in ONE kernel
    "backbone3d_four_kernel_layer": dict(ffn="kernel", backbone=True),   # ... out-projection + norm1 as a kernel of its own
    "backbone3d_two_kernel_ffn": dict(ffn="epilogue", backbone=True),   # ... FFN as FC + GELU, then FC + norms (round-2 mid form)
    "backbone3d_graph": dict(ffn="graph", backbone=True),    # ... in the reference graph's node structure
    "plugin_only": dict(ffn="off", backbone=False),          # plugins only; TensorRT-native layers stood in by fixed tensors
    "relaxed_tails": dict(ffn="layer", backbone=True, zero_tails=0),
    "backbone3d_postprocess": dict(ffn="layer", backbone=True, head=True),   # + CenterHead post-process graph + GPU NMS
    # ... with the head maps computed from the frame's own BEV map by a cuDNN stand-in of the 2-D backbone + CenterHead
    # convolutions (library code, random weights): raw points -> boxes after NMS
    "whole_pipeline": dict(ffn="layer", backbone=True, head="conv"),
}


class Slot:
    """One frame slot: buffers, captured graph, pinned host staging."""

    def __init__(self, pipeline_mod, cfg, weights, precision, cloud, seed, kind="backbone3d", share=None):
        import torch
        self.frame = pipeline_mod.HotPathFrame(cfg, weights, precision=precision, seed=seed, **FRAME_KINDS[kind])
        self.n = len(cloud)
        if share is not None:          # another leg over the same cloud: reuse the pinned staging buffers
            self.host_points, self.host_n = share.host_points, share.host_n
            self.host_boxes, self.host_valid = share.host_boxes, share.host_valid
            self.frame.points.copy_(share.frame.points)
            self.frame.points_size.copy_(share.frame.points_size)
        else:
            self.host_points = torch.from_numpy(cloud).pin_memory()
            self.host_n = torch.tensor([self.n], dtype=torch.int32).pin_memory()
            self.host_boxes = torch.empty(cfg.max_top_k, 9, dtype=torch.float32).pin_memory()
            self.host_valid = torch.empty(1, dtype=torch.int32).pin_memory()
            self.frame.load_points(cloud)
        self.graph = None

    def capture(self, stream):
        import torch
        with torch.cuda.stream(stream):
            if getattr(self.frame, "conv", None) is not None:
                self.frame.calibrate_head()       # random-weight head: heat-map sparsity of a trained one
            self.frame.run()                      # warm-up launch outside the capture
            stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=stream):
                self.frame.run()
        stream.synchronize()
        return self

    def enqueue_device(self):
        self.graph.replay()

    def enqueue_host(self):
        f = self.frame
        f.points[0, : self.n].copy_(self.host_points, non_blocking=True)
        f.points_size.copy_(self.host_n, non_blocking=True)
        self.graph.replay()
        self.host_boxes.copy_(f.boxes[0], non_blocking=True)
        self.host_valid.copy_(f.valid, non_blocking=True)


def run_steps(slots, streams, n_steps, host):
    """Runs n_steps steps; returns total device ms (events on a coordinating stream)."""
    import torch
    main = torch.cuda.current_stream()
    total = 0.0
    for _ in range(n_steps):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(main)
        for s in streams:
            s.wait_event(start)
        for i, slot in enumerate(slots):
            with torch.cuda.stream(streams[i % len(streams)]):
                slot.enqueue_host() if host else slot.enqueue_device()
        for s in streams:
            main.wait_stream(s)
        end.record(main)
        end.synchronize()
        total += start.elapsed_time(end)
    return total


def plugin_breakdown(slot, cfg, peaks, reps=5):
    """Per-launch device time of every kernel group of the HEADLINE frame (CUDA events around single eager launches on the
    slot's own buffers after one full run, L2 flushed before each repetition, median) + roofline numbers.
    calls_per_frame is the launch count in the backbone3d frame (FFN in its fused form)."""
    import torch
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    f = slot.frame
    assert f.backbone and f.ffn == "layer", "the breakdown describes the headline frame kind"
    f.run()
    torch.cuda.synchronize()
    V, Pc, P = int(f.vox.pillar_num[0]), int(f.vox.point_num[0]), slot.n
    W = [int(f.wp[i].win_num[0]) for i in (0, 1)]
    NS = [int(f.gs[i].set_num[0]) for i in (0, 1)]
    C, Fc, S = cfg.channel_num, cfg.ffn_channel_num, cfg.voxel_num_set
    F0, F1 = cfg.pfn_channels
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    flush_r = torch.zeros(64 * 1024 * 1024, dtype=torch.int32, device="cuda")

    def timed(fn):
        ts = []
        for _ in range(reps):
            flush.zero_()                       # L2 flush: 256 MB > 126 MB L2 ...
            flush_r.max()                       # ... then a 256 MB READ pass, so that the L2 is left full of CLEAN lines: after the
                                                # write alone the timed kernel pays for writing ~126 MB of dirty flush lines back
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    w, g, vox = f.w, f.w.glue, f.vox
    Vt = vox.pillar_num
    x = f.max_voxel[-1]                          # the VFE output: real pillar features
    res = {}
    # empty-kernel floor of this timing protocol (the latency-bound plugins are quoted against it): a GELU
    # launch over zero valid rows without tail fill -- a full-width grid whose CTAs exit at once, through the same C ABI
    tiny, none_valid = torch.zeros(8, 384, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
    tiny_out = torch.empty_like(tiny)
    res["_launch_floor"] = {"us": timed(lambda: capi.gelu(tiny, none_valid, out=tiny_out, zero_tails=0)), "bytes": 0,
                            "calls_per_frame": 0, "note": "empty launch under the same events + L2-flush protocol"}
    us = timed(lambda: vox(f.points, f.points_size))
    res["points2features"] = {"us": us, "bytes": 16 * P + 44 * Pc + 20 * V + 8, "calls_per_frame": 1}
    for i in (0, 1):
        us = timed(lambda: f.gs[i](f.wp[i].global_index, f.wp[i].coors_in_win, f.wp[i].voxel_num_in_win, f.wp[i].win_num))
        res[f"get_set_{i}"] = {"us": us, "bytes": 16 * V + 4 * W[i] + 2880 * NS[i] + 4, "calls_per_frame": 1}
        us = timed(lambda: f.wp[i](vox.coords, Vt))
        res[f"window_partition_{i}"] = {"us": us, "bytes": 16 * V + 16 * V + 20 * V + 4 * W[i], "calls_per_frame": 1,
                                        "scope": "next"}
    # VFE.  Headline: ONE kernel (dsvt_vfe_fused_launch: PFN 0, per-pillar max, concat, PFN 1, per-pillar max; bytes = point rows
    # in, one 768-byte row per pillar out).  Separate-launch form (legs.backbone3d_two_kernel_ffn / _graph): PFN layer 0
    # (streaming 10 -> 96), scatter-max, PFN layer 1 on [points | max] (tcgen05 linear), scatter-max -- timed on scratch buffers
    us = timed(lambda: capi.vfe_fused(g["pfn0"], g["pfn1"], vox.point_features[0], vox.point_index_in_voxel[0], Vt, vox.point_num,
                                      out=f.max_voxel[-1], workspace=f.vfe_ws))
    res["vfe_fused"] = {"us": us, "bytes": 40 * Pc + 4 * F1 * V + 8 * V, "flops": 2 * Pc * (10 * F0 + 2 * F0 * F1),
                        "mma_flops_issued": 3 * 2 * Pc * 2 * F0 * F1, "calls_per_frame": 1,
                        "scope": "next#3 (2 x TorchScatterMaxPlugin) + next#4 (PFN linears)", "launches": 2}
    Pf = cfg.max_points_num_voxel_filter
    sep = None if f.pfn0_out is not None else \
        [torch.empty(Pf, F0, device="cuda"), torch.empty(Pf, F1, device="cuda"), torch.empty(Pf, F0, device="cuda"),
         torch.empty(Pf, F1, device="cuda"), torch.empty(cfg.max_pillars_num, F0, device="cuda")]
    pfn0_out, pfn1_out = (f.pfn0_out, f.pfn1_out) if sep is None else (sep[0], sep[1])
    max_point = f.max_point if sep is None else [sep[2], sep[3]]
    max_voxel = f.max_voxel if sep is None else [sep[4], f.max_voxel[-1]]
    us = timed(lambda: g["pfn0"](vox.point_features[0], vox.point_num, activation=2, out=pfn0_out, zero_tails=0))
    res["pfn0_linear_bn_relu"] = {"us": us, "bytes": 4 * Pc * (10 + F0), "calls_per_frame": 0, "scope": "next#4"}
    capi.torch_scatter_max(pfn0_out, vox.point_index_in_voxel[0], vox.point_num_in_voxel[0], Vt, vox.point_num,
                           max_point=max_point[0], max_voxel=max_voxel[0])
    us = timed(lambda: g["pfn1"].rows_concat(pfn0_out, max_point[0], vox.point_num, activation=2, out=pfn1_out, zero_tails=0))
    res["pfn1_linear_bn_relu"] = {"us": us, "bytes": 4 * Pc * (2 * F0 + F1), "flops": 2 * Pc * 2 * F0 * F1,
                                  "calls_per_frame": 0, "scope": "next#4"}
    for k, (fch, src) in enumerate(((F0, pfn0_out), (F1, pfn1_out))):
        us = timed(lambda: capi.torch_scatter_max(src, vox.point_index_in_voxel[0], vox.point_num_in_voxel[0], Vt,
                                                  vox.point_num, max_point=max_point[k], max_voxel=max_voxel[k]))
        res[f"torch_scatter_max_{fch}"] = {"us": us, "bytes": 4 * fch * (2 * Pc + V) + 4 * (Pc + V), "calls_per_frame": 0,
                                           "scope": "next", "contract_bytes": 4 * fch * (Pc + Pf + cfg.max_pillars_num)}
    del sep, pfn0_out, pfn1_out, max_point, max_voxel
    torch.cuda.empty_cache()
    # position embedding.  Headline: the eight MLPs evaluated on the CELLS of a window (tables of 24 x 24 rows, one launch), looked
    # up by the QKV kernel; per-voxel forms (one launch per MLP / one launch for the eight) timed on scratch buffers
    pairs = [(b_, e_) for b_ in range(cfg.num_blocks) for e_ in (0, 1)]
    us = timed(lambda: capi.pos_embed_mlp_batch([g["pos"][b_][e_][0] for b_, e_ in pairs], [g["pos"][b_][e_][1] for b_, e_ in pairs],
                                                [f.pos_cells[e_] for b_, e_ in pairs], f.pos_rows,
                                                [f.pos_tab[b_][e_] for b_, e_ in pairs], zero_tails=0))
    ncell = int(f.pos_rows[0])
    res["pos_embed_tables_x8"] = {"us": us, "bytes": 8 * 4 * ncell * (2 + C), "flops": 8 * 2 * ncell * C * C, "calls_per_frame": 1,
                                  "scope": "next#4", "note": f"the eight position-embedding MLPs on the {ncell} cells of a window (one launch)"}
    pos_scratch = [torch.empty(cfg.max_pillars_num, C, device="cuda") for _ in range(8)]
    first, second = g["pos"][0][0]
    us = timed(lambda: capi.pos_embed_mlp(first, second, f.wp[0].coors_in_win_x_y[0], Vt, out=pos_scratch[0], zero_tails=0))
    res["pos_embed_mlp"] = {"us": us, "bytes": 4 * V * (2 + C), "flops": 2 * V * C * C, "calls_per_frame": 0, "scope": "next#4",
                            "note": "per voxel: Linear(2->192)+BN+ReLU generated inside the Linear(192->192) GEMM's producers; one MLP per launch"}
    us = timed(lambda: capi.pos_embed_mlp_batch([g["pos"][b_][e_][0] for b_, e_ in pairs], [g["pos"][b_][e_][1] for b_, e_ in pairs],
                                                [f.wp[e_].coors_in_win_x_y[0] for b_, e_ in pairs], Vt, pos_scratch, zero_tails=0))
    del pos_scratch
    res["pos_embed_mlp_x8"] = {"us": us, "bytes": 8 * 4 * V * (2 + C), "flops": 8 * 2 * V * C * C, "calls_per_frame": 0,
                               "scope": "next#4", "note": "per voxel: the frame's eight MLPs in one launch (backbone3d_four_kernel_layer leg)"}
    us = timed(lambda: capi.map2bev(f.final, vox.coords[0], Vt, cfg.grid_x, cfg.grid_y, out=f.bev))
    res["map2bev"] = {"us": us, "bytes": 2 * 4 * C * V + 16 * V, "calls_per_frame": 1, "scope": "next",
                      "contract_bytes": 4 * C * (cfg.grid_x * cfg.grid_y + V)}
    us = timed(lambda: capi.gelu(f.gelu_out, Vt, out=f.ffn_h))
    res["gelu"] = {"us": us, "bytes": 2 * 4 * Fc * V, "calls_per_frame": 0,
                   "note": "GeluPlugin alone (the headline frame folds it into the first FFN linear's epilogue)"}
    us = timed(lambda: capi.layer_norm(f.attn_out, Vt, w.gamma[0], w.beta[0], cfg.layer_norm_eps, residual=x, out=f.src_b))
    res["layer_norm_residual"] = {"us": us, "bytes": 3 * 4 * C * V + 2 * 4 * C, "calls_per_frame": 0,
                                  "note": "LayerNormPlugin (+ the kSUM in front) alone; in the headline frame all 28 LayerNorms run "
                                          "in GEMM epilogues"}
    us = timed(lambda: capi.filter_box(cfg, *f.cand, boxes=f.boxes, valid=f.valid))
    res["filter_box"] = {"us": us, "bytes": 22000 + 18004, "calls_per_frame": 1}
    # post-process graph + rotated NMS (legs.backbone3d_postprocess; not part of the headline frame)
    pipeline_mod = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    hf = pipeline_mod.HotPathFrame(cfg, w, precision=f.precision, seed=0, ffn="off", backbone=False, head=True)
    HWc = cfg.grid_x * cfg.grid_y
    us = timed(lambda: hf.topk(*hf.head_maps))
    res["center_head_topk"] = {"us": us, "bytes": 4 * 10 * HWc + cfg.max_top_k * 13 * 4, "calls_per_frame": 0, "scope": "next#4",
                               "note": "sigmoid + two-stage TopK(500) + gathers + exp / atan of src/dsvt-ai-trt.cpp:1471-1691; 4 kernels"}
    capi.filter_box(cfg, *hf.topk.outputs, boxes=hf.boxes, valid=hf.valid)
    us = timed(lambda: hf.nms(hf.boxes, hf.valid))
    res["rotated_nms"] = {"us": us, "bytes": 2 * cfg.max_top_k * 36, "calls_per_frame": 0, "scope": "next#4",
                          "note": f"nms_cpu (include/helper.h:257-283) on the GPU: {int(hf.valid[0])} boxes in, {int(hf.nms.num[0])} out; 3 kernels"}
    del hf
    # FFN (src/dsvt-ai-trt.cpp:494-529) on the FP32-accurate tcgen05 linear kernel; the second linear carries the LayerNorms
    # behind the FFN (norm2, norm, and the block's residual norm on every second layer) in its epilogue
    fc1, fc2 = w.ffn[0]
    us = timed(lambda: fc1.rows(f.src, Vt, activation=1, out=f.gelu_out, zero_tails=0))
    res["ffn_linear1_gelu"] = {"us": us, "flops": 2 * C * Fc * V, "bytes": 4 * V * (C + Fc), "calls_per_frame": 0, "scope": "next#4",
                               "note": "two-kernel FFN (legs.backbone3d_two_kernel_ffn); the headline frame runs ffn_fused_norm*"}
    st2 = [(f.src, w.gamma[1], w.beta[1]), (x, w.gamma[2], w.beta[2])]
    st3 = st2 + [(x, w.gamma[3], w.beta[3])]
    us = timed(lambda: fc2.rows_norm(f.gelu_out, Vt, st2, cfg.layer_norm_eps, out=f.src_b))
    res["ffn_linear2_norm2"] = {"us": us, "flops": 2 * C * Fc * V, "bytes": 4 * V * (Fc + 2 * C + C), "calls_per_frame": 0,
                                "scope": "next#4 + 2 LayerNormPlugins"}
    us = timed(lambda: fc2.rows_norm(f.gelu_out, Vt, st3, cfg.layer_norm_eps, out=f.src_b))
    res["ffn_linear2_norm3"] = {"us": us, "flops": 2 * C * Fc * V, "bytes": 4 * V * (Fc + 3 * C + C), "calls_per_frame": 0,
                                "scope": "next#4 + 3 LayerNormPlugins"}
    # the FFN alone: FC 192->384 + GELU + FC 384->192 + the LayerNorm chain in ONE kernel (dsvt_ffn_fused_launch; the
    # backbone3d_four_kernel_layer leg); bytes = x in, residual rows in, y out -- the 384-wide hidden rows never reach memory
    for n_st, st in ((2, st2), (3, st3)):
        us = timed(lambda: fc1.ffn_norm(fc2, f.src, Vt, st, cfg.layer_norm_eps, out=f.src_b))
        res[f"ffn_fused_norm{n_st}"] = {"us": us, "flops": 4 * C * Fc * V, "mma_flops_issued": 3 * 4 * C * Fc * V,
                                        "bytes": 4 * V * (C + n_st * C + C), "calls_per_frame": 0,
                                        "scope": f"next#4 (both FFN linears + GeluPlugin) + {n_st} LayerNormPlugins"}
    # the headline's layer tail: attention out-projection + norm1 + FFN + the LayerNorm chain in ONE kernel
    # (dsvt_attention_tail_ffn_launch), timed on the workspace a QKV + core launch pair has just filled; bytes = core rows in,
    # x in, src out, y out (+ the block input for the third norm)
    gs0, plan0 = f.gs[0], f.plans[(0, 0)]
    capi.set_attention_fused(w.attn[0], x, None, gs0.global_index_in_set[0], gs0.mask_expand_0[0], gs0.set_num, Vt, axis=0,
                             out=f.src_b, precision=f.precision, workspace=f.attn_ws, plan=plan0, stages=3, pos_table=f.attn_pos(0, 0)[1])
    for n_st, st in ((2, st2), (3, st3)):
        us = timed(lambda: capi.attention_tail_ffn(w.attn[0], fc1, fc2, x, gs0.global_index_in_set[0], Vt, 0, plan0, f.attn_ws,
                                                   (w.gamma[0], w.beta[0], cfg.layer_norm_eps), st, cfg.layer_norm_eps,
                                                   src=f.src, out=f.src_b))
        res[f"attention_tail_ffn_norm{n_st}"] = {
            "us": us, "flops": 2 * C * C * V + 4 * C * Fc * V, "mma_flops_issued": 3 * (2 * C * C * V + 4 * C * Fc * V),
            "bytes": 4 * V * C * (4 + (n_st - 2)), "calls_per_frame": 4,
            "scope": f"a3 out-projection + LayerNormPlugin (norm1) + next#4 (FFN) + {n_st} LayerNormPlugins"}
    pipeline_prec = f.precision in (capi.DSVT_ATTN_FP32_TC, capi.DSVT_ATTN_FP16_GEMM)
    for i in (0, 1):
        gs = f.gs[i]
        plan = f.plans.get((i, 0)) if getattr(f, "plans", None) else None
        call = lambda stages=7: capi.set_attention_fused(
            w.attn[i], x, f.attn_pos(i, 0)[0], gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num, Vt, axis=0,
            out=f.src_b, precision=f.precision, workspace=f.attn_ws, plan=plan,
            norm=(x, w.gamma[0], w.beta[0], cfg.layer_norm_eps), stages=stages, pos_table=f.attn_pos(i, 0)[1])
        if plan is not None:      # one plan per (partition, axis) serves two layers: 4 plan builds per frame
            pus = timed(lambda: capi.set_attention_plan(gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num, 0,
                                                        cfg.max_pillars_num, cfg.num_heads, cfg.channel_num, out=plan))
            res[f"set_attention_plan_{i}"] = {"us": pus, "bytes": NS[i] * (S * 4 + 8 * S * 4) + 4 * V + NS[i] * S * 8,
                                              "calls_per_frame": 2}
        flops = 11612160 * NS[i] if S == 36 else None
        if pipeline_prec and plan is not None:
            # the headline frame launches the QKV projection and the per-set core here; the out-projection (+ norm1) runs inside
            # attention_tail_ffn_norm*.  One launch each (dsvt_set_attention_fused_stages_launch), CUDA events around
            a, b, c = timed(lambda: call(1)), timed(lambda: call(2)), timed(lambda: call(4))
            split = 3 if f.precision == capi.DSVT_ATTN_FP32_TC else 1
            res[f"set_attention_{i}"] = {"us": a + b, "flops": None if flops is None else flops - 2 * C * C * V,
                                         "bytes": 83088 * NS[i] + 590000, "calls_per_frame": 4,
                                         "note": "QKV projection + per-set core (the out-projection is part of attention_tail_ffn_norm*)"}
            res[f"set_attention_{i}"]["kernels"] = {
                "qkv_proj_gemm": {"us": round(a, 2), "flops": 2 * 3 * C * C * V, "mma_flops_issued": split * 2 * 3 * C * C * V,
                                  "bytes": 4 * V * (2 * C + 3 * C) + 3 * 2 * split * 2 * C * C},
                "attn_core": {"us": round(b, 2), "bytes": 4 * V * (3 * C + C) + NS[i] * (S * 4 + 8 * S * 4),
                              "flops": None}}
            res[f"set_attention_out_proj_norm1_{i}"] = {
                "us": c, "flops": 2 * C * C * V, "mma_flops_issued": split * 2 * C * C * V, "bytes": 4 * V * 3 * C + 2 * split * 2 * C * C,
                "calls_per_frame": 0, "note": "out-projection + norm1 as a launch of its own (backbone3d_four_kernel_layer leg)"}
        else:
            res[f"set_attention_{i}"] = {"us": timed(call), "flops": flops, "bytes": 83088 * NS[i] + 590000, "calls_per_frame": 4}
    for k, r in res.items():
        if r.get("bytes"):
            r["gbs"] = r["bytes"] / r["us"] * 1e-3
            r["hbm_frac"] = r["gbs"] / peaks["hbm_gbs"]
        if r.get("flops"):
            r["tflops"] = r["flops"] / r["us"] * 1e-6
            r["tensor_frac"] = r["tflops"] / peaks["bf16_tflops"]
        r["us"] = round(r["us"], 2)
    frame_us = sum(r["us"] * r["calls_per_frame"] for r in res.values())
    stats = {"points": P, "kept_points": Pc, "pillars": V, "windows": W, "sets": NS}
    return res, frame_us, stats


def cpu_baseline(cfg, cloud, stats):
    """CPU oracle port on ONE core: one frame, attention on a bounded sample of sets and scaled."""
    import numpy as np
    from oracle import cpu
    t0 = time.perf_counter()
    pts = np.zeros((cfg.max_points_num, 4), np.float32)
    pts[: len(cloud)] = cloud
    o = cpu.points2features(pts, len(cloud), cfg)
    V = o["pillar_num"]
    t_part, parts = 0.0, []
    for i in (0, 1):
        owp = cpu.window_partition(o["coords"], V, cfg, i)
        parts.append(cpu.get_set(owp["global_index"], owp["coors_in_win"], owp["voxel_num_in_win"], owp["win_num"], cfg, i))
    t_index = time.perf_counter() - t0
    rng = np.random.default_rng(0)
    C, Fc = cfg.channel_num, cfg.ffn_channel_num
    x = rng.standard_normal((cfg.max_pillars_num, C)).astype(np.float32)
    h = rng.standard_normal((cfg.max_pillars_num, Fc)).astype(np.float32)
    g, b = np.ones(C, np.float32), np.zeros(C, np.float32)
    t0 = time.perf_counter(); cpu.layer_norm(x, V, g, b, 0.0, residual=x); t_ln = time.perf_counter() - t0
    t0 = time.perf_counter(); cpu.gelu(h, V); t_gelu = time.perf_counter() - t0
    pfn = [rng.standard_normal((cfg.max_points_num_voxel_filter, fch), dtype=np.float32) for fch in cfg.pfn_channels]
    t0 = time.perf_counter()
    for pf in pfn:
        cpu.torch_scatter_max(pf, o["point_index_in_voxel"], o["point_num_in_voxel"], V)
    cpu.map2bev(x, o["coords"], V, cfg.grid_x, cfg.grid_y)
    t_glue = time.perf_counter() - t0
    sample = 48
    q = rng.standard_normal((sample, cfg.voxel_num_set, C)).astype(np.float32)
    mask = np.zeros((sample, cfg.num_heads, cfg.voxel_num_set), np.float32)
    w_in = (rng.standard_normal((3 * C, C)) * 0.06).astype(np.float32)
    w_out = (rng.standard_normal((C, C)) * 0.06).astype(np.float32)
    t0 = time.perf_counter()
    cpu.set_attention(q, q, q, mask, sample, w_in, np.zeros(3 * C, np.float32), w_out, np.zeros(C, np.float32))
    t_set = (time.perf_counter() - t0) / sample
    n_sets = sum(p["set_num"] for p in parts) * 4          # 4 attention calls per partition
    frame_s = t_index + t_glue + 28 * t_ln + 8 * t_gelu + t_set * n_sets
    return {"value": 1.0 / frame_s, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 frame of {len(cloud)} pts: voxelise+partition {t_index:.3f}s, LayerNorm {t_ln:.3f}s x28, "
                      f"GELU {t_gelu:.3f}s x8, scatter-max x2 + map2bev {t_glue:.3f}s measured in full; set attention measured on {sample} sets "
                      f"({t_set*1e3:.2f} ms/set) and scaled to {n_sets} sets",
            "host_cores_available": os.cpu_count()}


# ---------------------------------------------------------------------------------------------
def main():
    """Keeps stdout clean for the ONE JSON line: libraries (NCCL prints its version banner to stdout on the first
    collective) write to fd 1, so fd 1 points at stderr while the bench runs and the line goes to the real stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _main()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if isinstance(line, dict):
        print(json.dumps(line), flush=True)
    return 0


def workload_config(args, cfg, world, F, S):
    """The `config` object -- byte-identical in both arms (`--impl ours` / `--impl reference`)."""
    return {"workload": f"BASELINE.json configs[1]: {args.points}-pt synthetic ring-lidar clouds (SURVEY 8(d) generator, seed = "
                        f"global frame id), pillar {cfg.voxel_x:g}x{cfg.voxel_y:g} (grid {cfg.grid_x}), 4 DSVT blocks, set="
                        f"{cfg.voxel_num_set}, FP32; '3-D backbone frame' of SURVEY 8(d): voxeliser -> PFN -> scatter-max -> window "
                        "partition -> getSet -> 8 x (set attention, FFN, LayerNorms, position embedding) -> BEV map, + filterBoxByScore "
                        "on synthetic head candidates; the 2-D BEV backbone + CenterHead are NOT executed",
            "points_per_frame": args.points, "frames_per_step_per_gpu": F, "streams_per_gpu": S,
            "capacities": {"max_points": cfg.max_points_num, "max_pillars": cfg.max_pillars_num, "max_sets": cfg.max_win_num},
            "parallelism": f"frame-parallel x{world} (frame f -> rank f mod N)",
            "l2": "no explicit flush: each step touches F frames x ~0.6 GB of distinct buffers >> 126 MB L2"}


def timed_leg(slots, streams, args, world, host_too=True):
    """warm-up + K timed steps, device-resident and (optionally) through host buffers; max over ranks -> (dev_ms, e2e_ms)."""
    run_steps(slots, streams, args.warmup, host=False)
    barrier(world)
    dev_ms = run_steps(slots, streams, args.steps, host=False)
    barrier(world)
    dev_ms = max_over_ranks(dev_ms, world)
    e2e_ms = None
    if host_too:
        run_steps(slots, streams, max(1, args.warmup), host=True)
        barrier(world)
        e2e_ms = run_steps(slots, streams, args.steps, host=True)
        barrier(world)
        e2e_ms = max_over_ranks(e2e_ms, world)
    return dev_ms, e2e_ms


def config4_leg(args, cfg, world, rank, make_frame_slot, n_frames=64, points=180000):
    """BASELINE.json configs[3] as written: 64 frames x 180k points, frame f -> rank f mod N, ONE FRAME PER STREAM, the
    results gathered with NCCL -- all inside the timed region (H2D of every cloud, the frames, D2D packing of the
    [500,9] boxes + counts = 18 004 B per frame, one batched gather to rank 0).  Strong scaling: 64 frames in total."""
    import torch
    pkg = importlib.import_module("dsvt-ai-trt_b200")
    sharding = importlib.import_module("dsvt-ai-trt_b200.sharding")
    ids = sharding.frames_for_rank(n_frames, rank, world)
    streams = [torch.cuda.Stream() for _ in ids]
    slots = [make_frame_slot(pkg.synth.ring_lidar(points, seed=f), f).capture(st) for f, st in zip(ids, streams)]
    K = cfg.max_top_k
    packed = torch.empty(len(slots), K * 9 + 1, device="cuda")          # per frame: 500 x 9 f32 boxes + the count (as bits)
    main = torch.cuda.current_stream()

    def one_pass():
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(main)
        for i, (s, st) in enumerate(zip(slots, streams)):
            st.wait_event(start)
            with torch.cuda.stream(st):
                s.frame.points[0, : s.n].copy_(s.host_points, non_blocking=True)
                s.frame.points_size.copy_(s.host_n, non_blocking=True)
                s.graph.replay()
                packed[i, : K * 9].copy_(s.frame.boxes[0].reshape(-1), non_blocking=True)
                packed[i, K * 9:].copy_(s.frame.valid.view(torch.float32), non_blocking=True)
            main.wait_stream(st)
        gathered = sharding.gather_packed(packed, dst=0)                 # NCCL gather (identity at N = 1)
        end.record(main)
        end.synchronize()
        return start.elapsed_time(end), gathered

    for _ in range(max(1, args.warmup)):
        one_pass()
    barrier(world)
    total, gathered = 0.0, None
    for _ in range(args.steps):
        ms, gathered = one_pass()
        total += ms
    barrier(world)
    total = max_over_ranks(total, world)
    boxes_kept = None
    if gathered is not None:
        boxes_kept = int(gathered[:, K * 9].contiguous().view(torch.int32).sum())
    del slots
    torch.cuda.empty_cache()
    return {"value": round(n_frames * args.steps / (total * 1e-3), 2), "unit": UNIT, "scaling": "strong",
            "ms_per_pass": round(total / args.steps, 3), "frames": n_frames, "points_per_frame": points,
            "frames_this_rank": len(ids), "streams_this_rank": len(ids),
            "gather": {"backend": "nccl" if world > 1 else "none (1 rank)", "bytes_per_frame": (K * 9 + 1) * 4,
                       "inside_timed_region": True},
            "h2d_bytes_per_pass": n_frames * (points * 16 + 4), "boxes_gathered": boxes_kept,
            "workload": "BASELINE.json configs[3]: 64 synthetic Waymo-shape frames (180k points each, seeds 0-63), one frame per "
                        "stream, frame f -> rank f mod N, NCCL gather of the boxes; same frame kind as the headline"}


def load_ncu_traffic():
    """DRAM bytes per launch of the profiled kernels: profiles/ncu_traffic.json, written by tools/ncu_traffic.py from a
    committed `ncu --set full` capture of this bench's frame."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return {}
    try:
        return json.load(open(path))
    except Exception:
        return {}


def _main():
    args = parse()
    if args.impl == "reference":
        ref = importlib.import_module("bench_reference")
        return ref.main(args)          # a dict on rank 0, None elsewhere
    import numpy as np
    import torch
    world, rank, local = dist_setup(args)
    pkg = importlib.import_module("dsvt-ai-trt_b200")
    capi = importlib.import_module("dsvt-ai-trt_b200.capi")
    pipeline = importlib.import_module("dsvt-ai-trt_b200.pipeline")
    sharding = importlib.import_module("dsvt-ai-trt_b200.sharding")
    cfg = pkg.config.WAYMO
    peaks = load_peaks()
    # "fp32" = the FP32 configuration (attention tolerance 2e-5 vs the oracle) on the tensor-core path;
    # "fp32_cuda" = the same tolerance on the CUDA-core kernel (the exact-arithmetic yard-stick)
    precision = {"fp32": capi.DSVT_ATTN_FP32_TC, "fp32_cuda": capi.DSVT_ATTN_FP32, "fp32_tc": capi.DSVT_ATTN_FP32_TC,
                 "fp16": capi.DSVT_ATTN_FP16, "fp16_gemm": capi.DSVT_ATTN_FP16_GEMM}[args.precision]

    F, S = args.frames_per_step, max(1, min(args.streams, args.frames_per_step))
    weights = pipeline.FrameWeights(cfg, seed=0)
    streams = [torch.cuda.Stream() for _ in range(S)]
    clouds = [pkg.synth.ring_lidar(args.points, seed=sharding.global_frame_id(rank, world, i)) for i in range(F)]

    def make_slots(kind, prec, base=None):
        out = []
        for i in range(F):
            s = Slot(pipeline, cfg, weights, prec, clouds[i], sharding.global_frame_id(rank, world, i), kind=kind,
                     share=base[i] if base else None)
            out.append(s.capture(streams[i % S]))
        torch.cuda.synchronize()
        return out

    slots = make_slots("backbone3d", precision)
    launches_per_frame = slots[0].frame.launches_per_frame

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    dev_ms, e2e_ms = timed_leg(slots, streams, args, world)
    clocks = sampler.finish() if sampler else None

    # the "trivial NCCL result gather" of SURVEY 8(e) for the last step (the timed form is the config4 leg)
    boxes = torch.stack([s.frame.boxes[0] for s in slots])
    valid = torch.cat([s.frame.valid for s in slots])
    gathered = sharding.gather_results(boxes, valid, dst=0)

    def leg(kind, prec, note):
        ls = make_slots(kind, prec, base=slots)
        d, e = timed_leg(ls, streams, args, world)
        out = {"value": round(F * world * args.steps / (d * 1e-3), 2), "e2e": round(F * world * args.steps / (e * 1e-3), 2),
               "unit": UNIT, "launches_per_frame": int(ls[0].frame.launches_per_frame), "note": note}
        del ls
        torch.cuda.empty_cache()
        return out

    def single_stream_frame_seconds(kind, prec, reps=20):
        """Latency form: ONE frame slot, host cloud in -> host boxes out, nothing else in flight (median over reps)."""
        sl = Slot(pipeline, cfg, weights, prec, clouds[0], sharding.global_frame_id(rank, world, 0), kind=kind,
                  share=slots[0]).capture(streams[0])
        ts = []
        with torch.cuda.stream(streams[0]):
            for r in range(reps + 3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                sl.enqueue_host()
                b.record()
                b.synchronize()
                if r >= 3:
                    ts.append(a.elapsed_time(b) * 1e-3)
        ts.sort()
        del sl
        torch.cuda.empty_cache()
        return round(ts[len(ts) // 2], 6)

    legs = {}
    if not args.no_legs and args.precision == "fp32":
        legs["plugin_only"] = leg("plugin_only", precision,
                                  "round 1's headline frame: the reference's ten plugins + the fused set attention; the TensorRT-"
                                  "native layers (PFN, position embedding, FFN linears) are NOT executed, fixed tensors stand in")
        legs["backbone3d_four_kernel_layer"] = leg("backbone3d_four_kernel_layer", precision,
                                                   "the headline's data flow with the attention's out-projection + norm1 as a kernel "
                                                   "of its own in front of the fused FFN kernel (4 kernels per encoder layer)")
        legs["backbone3d_two_kernel_ffn"] = leg("backbone3d_two_kernel_ffn", precision,
                                                "the headline's data flow with every FFN as two kernels (FC + GELU epilogue, then FC + "
                                                "LayerNorm-chain epilogue): the 384-wide hidden rows go through memory")
        legs["backbone3d_graph"] = leg("backbone3d_graph", precision,
                                       "the headline's data flow in the reference graph's node structure: FC -> GeluPlugin -> FC "
                                       "(the headline folds the GELU and the residual add into the linears' epilogues)")
        legs["fp16_config"] = leg("backbone3d", capi.DSVT_ATTN_FP16,
                                  "BASELINE.json configs[2]: same frames, set attention on the single fused FP16 tcgen05 kernel "
                                  "(QK^T / PV on tensor cores, FP32 accumulate), tolerance 1e-2")
        legs["backbone3d_postprocess"] = leg("backbone3d_postprocess", precision,
                                             "the headline frame + the CenterHead post-process graph (sigmoid / TopK / gathers / atan, "
                                             "src/dsvt-ai-trt.cpp:1471-1691) on synthetic head maps feeding FilterBoxByScorePlugin, + the "
                                             "rotated NMS the reference runs on the HOST (include/helper.h:257-283) as CUDA kernels")
        wp = leg("whole_pipeline", precision,
                 "raw points -> boxes after NMS: the backbone3d_postprocess leg with the head maps computed from the frame's own BEV "
                 "map by a cuDNN (PyTorch, BF16, channels-last, random weights) STAND-IN of the reference's TensorRT-native 2-D BEV "
                 "backbone + CenterHead convolutions (src/dsvt-ai-trt.cpp:1137-1468) -- library code, out of SURVEY 8's scope, no "
                 "kernel of this repo; launches_per_frame counts this repo's kernels only.  Comparable in scope to the reference "
                 "README's ~0.7 s per frame (FP32 TensorRT, RTX 3080-class, one 300k-point nuScenes cloud): seconds_per_frame below")
        wp["seconds_per_frame_single_stream"] = single_stream_frame_seconds("whole_pipeline", precision)
        legs["whole_pipeline"] = wp
        legs["relaxed_tails"] = leg("relaxed_tails", precision,
                                    "NOT the reference's contract: every plugin launched with zero_tails = 0 (rows beyond the "
                                    "valid counts left untouched; no consumer reads them) -- what the contract's zero tails cost")
    cfg4 = None
    if not args.no_legs and not args.no_config4 and args.precision == "fp32":
        del slots[1:]                    # frame 0's slot stays for the breakdown; the leg needs the memory of 64 slots at N = 1
        torch.cuda.empty_cache()
        cfg4 = config4_leg(args, cfg, world, rank,
                           lambda cloud, f: Slot(pipeline, cfg, weights, precision, cloud, f, kind="backbone3d"))

    if rank != 0:
        return None
    frames = F * world * args.steps
    value = frames / (dev_ms * 1e-3)
    e2e = frames / (e2e_ms * 1e-3)
    plugins, frame_us, stats = plugin_breakdown(slots[0], cfg, peaks)
    roof = roofline_block(plugins, frame_us, peaks)
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision.startswith("fp32") else args.precision,
        "dtype_note": {"fp32": "plugin I/O and all non-GEMM arithmetic FP32; projections / linears on tcgen05 with FP16 hi+lo split "
                               "operands (3 MMAs per product, FP32 accumulate in TMEM) held to the FP32 tolerance (2e-5 vs the oracle)",
                       "fp32_cuda": "all arithmetic on the FP32 CUDA cores"}.get(args.precision),
        "data": "synthetic",
        "config": workload_config(args, cfg, world, F, S),
        "frame_kind": "backbone3d (every layer of the reference's 3-D backbone as one data flow, src/dsvt-ai-trt.cpp:571-1128)",
        "frame_stats": stats,
        "e2e": {"value": round(e2e, 2), "unit": UNIT, "h2d_bytes_per_step": world * F * (args.points * 16 + 4),
                "d2h_bytes_per_step": world * F * (cfg.max_top_k * 9 * 4 + 4), "ms_per_step": round(e2e_ms / args.steps, 4),
                "bytes_note": "whole job: all ranks' pinned-host clouds in, boxes + counts out, per step"},
        "gpu_launches": int(launches_per_frame * F * args.steps * 2),
        "launches_per_frame": int(launches_per_frame),
        "clocks": clocks,
        "roofline": roof,
        "plugins": plugins,
        "frame_us_sum_of_plugins": round(frame_us, 1),
        "legs": legs,
        "config4_64x180k": cfg4,
        "gathered_boxes": None if gathered is None else int(gathered[1].sum()),
    }
    if not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(cfg, clouds[0], stats)
        except Exception as e:   # the checker must never take the bench down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {e}"}
    return line


def roofline_block(plugins, frame_us, peaks):
    """The dominant KERNEL by share of the frame: algorithmic bytes / flops (SURVEY 8(d) figures x units per launch) over
    its live CUDA-event duration, against MEASURED_PEAKS.json."""
    # per-KERNEL view: a plugin that is one kernel counts as such; the GEMM-pipeline attention contributes its kernels
    kernels = {}
    for k, r in plugins.items():
        if not r["calls_per_frame"]:
            continue
        if "kernels" in r:      # the two window partitions run the SAME kernels: call-weighted mean
            for kn, kr in r["kernels"].items():
                a = kernels.setdefault(f"set_attention.{kn}", {"us": 0.0, "calls_per_frame": 0, "bytes": 0, "flops": 0,
                                                               "mma_flops_issued": 0})
                c = r["calls_per_frame"]
                a["us"] = (a["us"] * a["calls_per_frame"] + kr["us"] * c) / (a["calls_per_frame"] + c)
                for f_ in ("bytes", "flops", "mma_flops_issued"):
                    a[f_] = (a[f_] * a["calls_per_frame"] + (kr.get(f_) or 0) * c) / (a["calls_per_frame"] + c)
                a["calls_per_frame"] += c
        else:
            kernels[k] = r
    # the layer-tail kernel runs with two or three LayerNorm stages behind the FFN: ONE kernel, call-weighted mean of the two
    for base in ("attention_tail_ffn", "ffn_fused"):
        parts = [kernels.pop(f"{base}_norm{n}") for n in (2, 3) if f"{base}_norm{n}" in kernels]
        if parts:
            c = sum(p_["calls_per_frame"] for p_ in parts)
            kernels[base] = {"calls_per_frame": c,
                             **{f_: sum((p_.get(f_) or 0) * p_["calls_per_frame"] for p_ in parts) / c
                                for f_ in ("us", "bytes", "flops", "mma_flops_issued")}}
    dom_key = max(kernels, key=lambda k: kernels[k]["us"] * kernels[k]["calls_per_frame"])
    dom = kernels[dom_key]
    # which roof bounds it: arithmetic intensity of the work the kernel actually issues (flops / algorithmic byte) against
    # the machine's ridge (measured bf16 TFLOP/s / measured HBM GB/s)
    ridge = peaks["bf16_tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    issued = dom.get("mma_flops_issued") or dom.get("flops") or 0
    intensity = issued / dom["bytes"] if dom.get("bytes") else float("inf")
    if dom.get("flops") and intensity >= ridge:
        tf = (dom.get("mma_flops_issued") or dom["flops"]) / dom["us"] * 1e-6
        roof = {"kernel": dom_key, "bound": "tensor", "achieved": round(tf, 3), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": round(tf / peaks["bf16_tflops"], 5)}
    else:
        gbs = dom["bytes"] / dom["us"] * 1e-3
        roof = {"kernel": dom_key, "bound": "hbm", "achieved": round(gbs, 2), "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": round(gbs / peaks["hbm_gbs"], 5)}
    if dom.get("flops"):
        roof["flop_per_byte"] = round(intensity, 1)
        roof["ridge_flop_per_byte"] = round(ridge, 1)
        roof["algorithmic_tflops"] = round(dom["flops"] / dom["us"] * 1e-6, 2)
        if dom.get("mma_flops_issued"):
            roof["mma_tflops_issued"] = round(dom["mma_flops_issued"] / dom["us"] * 1e-6, 2)
    roof["peak_source"] = peaks["source"] + " (burst cuBLAS bf16 / copy bandwidth, MEASURED_PEAKS.json)"
    t = load_ncu_traffic().get(dom_key)
    roof["traffic"] = t.get("dram_bytes_per_launch") if t else None
    roof["traffic_source"] = t.get("source") if t else None
    roof["algorithmic_bytes"] = dom.get("bytes")
    roof["us"] = round(dom["us"], 2)
    roof["share_of_frame"] = round(dom["us"] * dom["calls_per_frame"] / frame_us, 3)
    roof["timing"] = "CUDA events around the kernel on its launch stream, L2 flushed (256 MB written, then 256 MB read: cold and clean) before each repetition, bench frame 0"
    attn_us = sum(plugins[k]["us"] * plugins[k]["calls_per_frame"] for k in plugins if k.startswith("set_attention_") and "plan" not in k)
    attn_flops = sum((plugins[k].get("flops") or 0) * plugins[k]["calls_per_frame"] for k in plugins if k.startswith("set_attention_"))
    if attn_us:
        roof["set_attention_plugin"] = {"us_per_frame": round(attn_us, 1), "share_of_frame": round(attn_us / frame_us, 3),
                                        "algorithmic_tflops": round(attn_flops / attn_us * 1e-6, 2),
                                        "frac_of_bf16_peak": round(attn_flops / attn_us * 1e-6 / peaks["bf16_tflops"], 5),
                                        "note": "flops = 11 612 160 x sets (SURVEY 8d: all 36 slots of every set)"}
    roof["kernels"] = {k: {"us": round(v["us"], 2), "calls_per_frame": v["calls_per_frame"],
                           "share_of_frame": round(v["us"] * v["calls_per_frame"] / frame_us, 3),
                           "hbm_frac": round(v["bytes"] / v["us"] * 1e-3 / peaks["hbm_gbs"], 4) if v.get("bytes") else None}
                       for k, v in kernels.items()}
    return roof


if __name__ == "__main__":
    rc = main()
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    sys.exit(rc)
