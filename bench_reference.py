"""bench.py --impl reference: the reference's OWN implementation of the hot path, timed on the same box.

The reference (jingyue202205/DSVT-AI-TRT) has no CPU compute path: its hot path is CUDA plugins + TensorRT
layers.  This arm therefore runs the reference's plugin sources, compiled UNMODIFIED for
sm_100a into oracle/_ref/waymo/ (capacities raised through oracle/ref_config_waymo/params.h, the reference's
own configuration mechanism), through the same TensorRT-style C harness, in the reference's graph order and
with the reference's tensor plumbing: GetValueByIndex -> MHA -> MapSetFeature2Voxel, separate elementwise
residual adds, every plugin's own full-capacity memsets.  multHeadAttention() executes inside closed-source
TensorRT in the reference; it is stood in by torch.nn.functional.multi_head_attention_forward (PyTorch eager,
FP32, over all max_win_num padded sets exactly like the reference graph).  None of this repo's kernels run on
this path.  If oracle/_ref is missing or the reference kernels fault, the arm falls back to timing the CPU
oracle port (kind "port").
"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_DIR = os.path.join(ROOT, "oracle", "_ref", "waymo")


class ReferenceFrame:
    def __init__(self, plg, cfg, cloud, seed, backbone=False):
        self.backbone = backbone
        import numpy as np
        import torch
        import torch.nn.functional as Fn
        self.torch, self.Fn, self.cfg = torch, Fn, cfg
        lib = lambda stem: plg.PluginLibrary(os.path.join(REF_DIR, f"libref_{stem}.so"))
        rng = np.random.default_rng(0)
        C, F, mp, mw, S = cfg.channel_num, cfg.ffn_channel_num, cfg.max_pillars_num, cfg.max_win_num, cfg.voxel_num_set
        self.vox = plg.add_voxel_generator(lib("points2Features"), cfg.max_points_num, cfg.max_points_num_voxel_filter,
                                           mp, 4, 10, cfg.max_num_points_per_voxel, cfg.x_min, cfg.x_max, cfg.y_min,
                                           cfg.y_max, cfg.z_min, cfg.z_max, cfg.voxel_x, cfg.voxel_y, cfg.voxel_z,
                                           cfg.grid_x, cfg.grid_y, cfg.grid_z)
        wl, gl = lib("windowPartition"), lib("getSet")
        self.wp = [plg.add_window_partition(wl, mw, cfg.max_voxel_num_per_win, (cfg.grid_x, cfg.grid_y, cfg.grid_z),
                                            cfg.win_shapes[i], cfg.shifts[i]) for i in (0, 1)]
        self.gs = [plg.add_get_set_op(gl, mw, cfg.max_voxel_num_per_win, S, cfg.win_shapes[i]) for i in (0, 1)]
        gv, ms = lib("getValueByIndex"), lib("mapSetFeature2voxel")
        self.gather = [plg.add_get_value_by_index_op(gv, mw, S, C, a) for a in (0, 1)]
        self.scatter = [plg.add_map_set_feature2voxel_op(ms, mw, S, C, a, mp) for a in (0, 1)]
        self.gelu = plg.add_gelu_op(lib("gelu"), mp, F)
        sml = lib("torchScatterMax")
        self.smax = [plg.add_torch_scatter_max(sml, cfg.max_points_num_voxel_filter, mp, f) for f in cfg.pfn_channels]
        self.m2b = plg.add_map_2_bev_op(lib("map2bev"), mp, C, cfg.grid_x, cfg.grid_y)
        ll = lib("layerNorm")
        self.ln = [plg.add_layer_norm_op(ll, mp, C, (1.0 + 0.1 * rng.standard_normal(C)).astype(np.float32),
                                         (0.1 * rng.standard_normal(C)).astype(np.float32))
                   for _ in range(cfg.num_blocks * 7)]
        self.fb = plg.add_filter_box_by_score_op(lib("filterBoxByScore"), cfg.max_top_k, cfg.x_min, cfg.x_max, cfg.y_min,
                                                 cfg.y_max, cfg.z_min, cfg.z_max, cfg.voxel_x, cfg.voxel_y, cfg.voxel_z,
                                                 cfg.score_threshold)
        dev = "cuda"
        g = torch.Generator().manual_seed(seed)
        self.n = len(cloud)
        self.points = torch.zeros(1, cfg.max_points_num, 4, device=dev)
        self.points[0, : self.n] = torch.from_numpy(cloud).to(dev)
        self.points_size = torch.tensor([self.n], dtype=torch.int32, device=dev)
        self.host_points = torch.from_numpy(cloud).pin_memory()
        self.host_boxes = torch.empty(cfg.max_top_k, 9).pin_memory()
        self.host_valid = torch.empty(1, dtype=torch.int32).pin_memory()
        g2 = torch.Generator().manual_seed(1)
        self.pfn_out = [torch.randn(1, cfg.max_points_num_voxel_filter, f, generator=g2).to(dev) for f in cfg.pfn_channels]
        self.x0 = torch.randn(1, mp, C, generator=g).to(dev)
        self.pos = [[torch.randn(1, mp, C, generator=g).mul_(0.5).to(dev) for _ in range(2)] for _ in range(cfg.num_blocks)]
        self.ffn_hidden = torch.randn(1, mp, F, generator=g).to(dev)
        self.ffn_out = torch.randn(1, mp, C, generator=g).mul_(0.5).to(dev)
        self.attn_w = [((torch.randn(3 * C, C, generator=g) * 0.06).to(dev), (torch.randn(3 * C, generator=g) * 0.02).to(dev),
                        (torch.randn(C, C, generator=g) * 0.06).to(dev), (torch.randn(C, generator=g) * 0.02).to(dev))
                       for _ in range(cfg.num_blocks * 2)]
        synth = importlib.import_module("dsvt-ai-trt_b200.synth")
        sc, cl, xs, ys, ce, cz, an, dm = synth.head_candidates(cfg.max_top_k, seed)
        K = cfg.max_top_k

        def carve(a, width):   # the reference reads 12 candidates past the end (SURVEY A-10): back them with zeros
            buf = torch.zeros(512 * width, dtype=torch.from_numpy(a).dtype, device=dev)
            buf[: K * width] = torch.from_numpy(a).to(dev).flatten()
            return buf[: K * width]
        self.cand = [carve(sc, 1).view(1, K), carve(cl, 1).view(1, K), carve(xs, 1).view(1, K), carve(ys, 1).view(1, K),
                     carve(ce, 2).view(1, 1, K, 2), carve(cz, 1).view(1, 1, K, 1), carve(an, 1).view(1, 1, K, 1),
                     carve(dm, 3).view(1, 1, K, 3)]
        if backbone:
            # the TensorRT-native layers of the 3-D backbone (FullyConnected + Scale + ReLU, src/dsvt-ai-trt.cpp:268-286,
            # :461-529), stood in by PyTorch eager (cuBLAS) over the full static shapes like the engine's layers
            F0, F1 = cfg.pfn_channels
            r = lambda *sh, k=1.0: (torch.randn(*sh, generator=g) * k).to(dev)
            self.pfn = [(r(F0, 10, k=0.05), r(F0, k=0.1) + 0.5, r(F0, k=0.1)), (r(F1, 2 * F0, k=0.07), r(F1, k=0.1) + 0.5, r(F1, k=0.1))]
            self.pos_w = [[(r(C, 2, k=0.3), r(C, k=0.1) + 0.5, r(C, k=0.1), r(C, C, k=0.07), r(C, k=0.02)) for _ in range(2)]
                          for _ in range(cfg.num_blocks)]
            self.ffn_w = [(r(F, C, k=0.07), r(F, k=0.02), r(C, F, k=0.05), r(C, k=0.02)) for _ in range(cfg.num_blocks * 2)]
        self.out = {}          # persistent output tensors per plugin call site (static addresses for graph capture)
        self.graph = None
        self.boxes = None

    def call(self, key, plugin, inputs):
        outs = plugin.enqueue(inputs, outputs=self.out.get(key))
        self.out[key] = outs
        return outs

    def mha(self, q, k, v, mask, w):
        """multHeadAttention() (src/dsvt-ai-trt.cpp:288-458) restated with PyTorch eager over ALL padded sets."""
        cfg, Fn = self.cfg, self.Fn
        C, H = cfg.channel_num, cfg.num_heads
        kpm = mask[0, :, 0, :] < 0                                   # [sets, S] key padding mask
        out, _ = Fn.multi_head_attention_forward(
            q[0].transpose(0, 1), k[0].transpose(0, 1), v[0].transpose(0, 1), C, H, w[0], w[1], None, None, False, 0.0,
            w[2], w[3], training=False, key_padding_mask=kpm, need_weights=False)
        return out.transpose(0, 1).contiguous()[None]

    def run(self):
        cfg, torch = self.cfg, self.torch
        Fn = self.Fn
        vo = self.call("vox", self.vox, [self.points, self.points_size])
        V = vo[4]
        x0 = self.x0
        if self.backbone:
            (w0, s0, t0), (w1, s1, t1) = self.pfn
            h0 = torch.relu(Fn.linear(vo[0], w0) * s0 + t0)                                 # PFN layer 0  (:577)
            sm0 = self.call("sm0", self.smax[0], [h0, vo[1], vo[3], V])
            h1 = torch.relu(Fn.linear(torch.cat([h0, sm0[0]], dim=2), w1) * s1 + t1)      # concat + PFN layer 1 (:583-587)
            x0 = self.call("sm1", self.smax[1], [h1, vo[1], vo[3], V])[1]
        else:
            for k in range(len(self.smax)):
                self.call(f"sm{k}", self.smax[k], [self.pfn_out[k], vo[1], vo[3], V])
        parts, wps = [], []
        for i in (0, 1):
            w = self.call(f"wp{i}", self.wp[i], [vo[2], V])
            wps.append(w)
            parts.append(self.call(f"gs{i}", self.gs[i], w[:4]))
        pos = self.pos
        if self.backbone:                                                                   # 8 position-embedding MLPs (:603-637)
            pos = [[Fn.linear(torch.relu(Fn.linear(wps[enc][5], a) * sc + sh), b2, bias2)
                    for enc, (a, sc, sh, b2, bias2) in enumerate(row)] for row in self.pos_w]
        x, ln = x0, 0
        for blk in range(cfg.num_blocks):
            gs = parts[blk % 2]
            x_in = x
            for enc in (0, 1):
                q, k, v = self.call(f"gv{enc}", self.gather[enc], [x, pos[blk][enc], gs[0], gs[2]])
                a = self.mha(q, k, v, gs[3], self.attn_w[blk * 2 + enc])
                y = self.call(f"ms{enc}", self.scatter[enc], [a, gs[0], gs[2]])[0]
                src = self.call(f"ln{ln}", self.ln[ln], [y + x, V])[0]; ln += 1
                if self.backbone:                                                           # FFN (:494-529)
                    f1, fb1, f2, fb2 = self.ffn_w[blk * 2 + enc]
                    hid = self.call("ge", self.gelu, [Fn.linear(src, f1, fb1), V])[0]
                    ffn_out = Fn.linear(hid, f2, fb2)
                else:
                    self.call("ge", self.gelu, [self.ffn_hidden, V])
                    ffn_out = self.ffn_out
                src = self.call(f"ln{ln}", self.ln[ln], [src + ffn_out, V])[0]; ln += 1
                x = self.call(f"ln{ln}", self.ln[ln], [src + x, V])[0]; ln += 1
            x = self.call(f"ln{ln}", self.ln[ln], [x + x_in, V])[0]; ln += 1
        self.call("m2b", self.m2b, [x, vo[2], V])
        self.boxes, self.valid = self.call("fb", self.fb, self.cand)
        return self

    def capture(self, stream):
        torch = self.torch
        with torch.cuda.stream(stream):
            self.run()
            stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=stream):
                self.run()
        stream.synchronize()

    def enqueue_device(self):
        self.graph.replay()

    def enqueue_host(self):
        self.points[0, : self.n].copy_(self.host_points, non_blocking=True)
        self.graph.replay()
        self.host_boxes.copy_(self.boxes[0], non_blocking=True)
        self.host_valid.copy_(self.valid, non_blocking=True)


def cpu_port_arm(args, cfg, pkg):
    """Fallback: the CPU oracle port timed on the host (one bounded frame sample per step)."""
    bench = importlib.import_module("bench")
    cloud = pkg.synth.ring_lidar(args.points, seed=0)
    vals = []
    for _ in range(max(1, min(args.steps, 2))):
        vals.append(bench.cpu_baseline(cfg, cloud, None))
    best = max(vals, key=lambda v: v["value"])
    return best


def main(args):
    """Returns the JSON line (a dict) on rank 0, None on the other ranks; bench.main() prints it.  The reference's
    creators print their fields to stdout (e.g. getSet.cu:829): bench.main() points fd 1 at stderr while this runs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None       # rank 0 alone runs the reference arm
    return _run(args)


def _run(args):
    import torch
    bench = importlib.import_module("bench")
    pkg = importlib.import_module("dsvt-ai-trt_b200")
    cfg = pkg.config.WAYMO
    line = {"impl": "reference", "metric": bench.METRIC, "unit": bench.UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}
    F, S = args.frames_per_step, max(1, min(args.streams, args.frames_per_step))
    try:
        if not (torch.cuda.is_available() and os.path.exists(os.path.join(REF_DIR, "libref_getSet.so"))):
            raise RuntimeError("oracle/_ref/waymo not available")
        torch.cuda.set_device(0)
        plg = importlib.import_module("dsvt-ai-trt_b200.plugins")
        streams = [torch.cuda.Stream() for _ in range(S)]
        slots = []
        for i in range(F):
            fr = ReferenceFrame(plg, cfg, pkg.synth.ring_lidar(args.points, seed=i), i)
            fr.capture(streams[i % S])
            slots.append(fr)
        torch.cuda.synchronize()
        bench.run_steps(slots, streams, args.warmup, host=False)
        torch.cuda.synchronize()
        dev_ms = bench.run_steps(slots, streams, args.steps, host=False)
        bench.run_steps(slots, streams, max(1, args.warmup), host=True)
        e2e_ms = bench.run_steps(slots, streams, args.steps, host=True)
        torch.cuda.synchronize()
        frames = F * args.steps
        value, e2e = frames / (dev_ms * 1e-3), frames / (e2e_ms * 1e-3)
        line.update({
            "value": round(value, 3), "ms_per_step": round(dev_ms / args.steps, 3),
            "config": {"workload": f"same plugin sequence and clouds as the 'ours' arm ({args.points}-pt ring-lidar, "
                                   f"capacities {cfg.max_points_num}/{cfg.max_pillars_num}/{cfg.max_win_num}); reference "
                                   "plugin sources compiled unmodified for sm_100a (oracle/_ref/waymo), MHA stood in by "
                                   "PyTorch eager over all padded sets, residual adds by torch, CUDA-graph replay",
                       "frames_per_step_per_gpu": F, "streams_per_gpu": S},
            "cpu_baseline": {"value": round(value, 3), "unit": bench.UNIT, "cores": 0, "kind": "reference",
                             "sample": "the reference's own CUDA kernels on the B200 (it has no CPU compute path); "
                                       "host cores only launch"},
            "e2e": {"value": round(e2e, 3), "unit": bench.UNIT, "h2d_bytes_per_step": F * args.points * 16,
                    "d2h_bytes_per_step": F * (cfg.max_top_k * 36 + 4)},
        })
        if not getattr(args, "no_ffn_leg", False):
            # the complete 3-D backbone (our arm's ffn_in_frame.backbone3d): + PFN, position-embedding and FFN linears as
            # PyTorch eager / cuBLAS with TF32 allowed (TensorRT's default on this class of GPU), full static shapes
            try:
                del slots
                torch.cuda.empty_cache()
                torch.backends.cuda.matmul.allow_tf32 = True
                slots_b = []
                for i in range(F):
                    fr = ReferenceFrame(plg, cfg, pkg.synth.ring_lidar(args.points, seed=i), i, backbone=True)
                    fr.capture(streams[i % S])
                    slots_b.append(fr)
                torch.cuda.synchronize()
                bench.run_steps(slots_b, streams, args.warmup, host=False)
                ms_b = bench.run_steps(slots_b, streams, args.steps, host=False)
                line["backbone3d"] = {"value": round(frames / (ms_b * 1e-3), 3), "unit": bench.UNIT,
                                      "note": "reference plugins + PyTorch-eager (cuBLAS, TF32 allowed) stand-ins for every "
                                              "TensorRT-native layer of the 3-D backbone, over the engine's full static shapes; TF32 also "
                                              "speeds up the MHA stand-in, which is why this leg can beat the FP32 plugin-only frame"}
            except Exception as exc2:
                line["backbone3d"] = {"value": None, "note": f"failed: {str(exc2)[:200]}"}
    except Exception as exc:     # reference kernels unavailable / faulted -> CPU oracle port
        cb = cpu_port_arm(args, cfg, pkg)
        line.update({
            "value": round(cb["value"], 5), "ms_per_step": round(1e3 / cb["value"], 1),
            "config": {"workload": f"CPU oracle port, 1 frame of {args.points} pts per step (bounded sample, see cpu_baseline)",
                       "fallback_reason": str(exc)[:200]},
            "cpu_baseline": cb,
            "e2e": {"value": round(cb["value"], 5), "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
    return line
