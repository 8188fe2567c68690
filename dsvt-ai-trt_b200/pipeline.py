"""One "hot-path frame": the plugin sequence of SURVEY.md section 8(a) in the reference's graph order
(src/dsvt-ai-trt.cpp:571-1120, :1684) for ONE point cloud, every launch going through the C ABI.

    a1 Points2Features -> [PFN linear 10->96: glue] -> TorchScatterMax(96) -> [PFN linear 192->192: glue]
                       -> TorchScatterMax(192) -> WindowPartition x2 -> a2 GetSet x2
    4 DSVT blocks x 2 encoders: a3 set attention (gather/scatter fused) -> a5 LayerNorm(y + x)
                                -> [FFN linear 192->384: TensorRT-native glue, NOT part of the hot path]
                                -> a4 GELU -> [FFN linear 384->192: glue] -> a5 LayerNorm -> a5 LayerNorm
                 + a5 LayerNorm per block                                   (7 LayerNorms per block, 28 per frame)
    Map2Bev (dense BEV map for the 2-D backbone)
    a6 FilterBoxByScore on the CenterHead's top-500 candidates

The TensorRT-native layers between the plugins (PFN, pos-embed MLPs, FFN linears, BEV backbone, head; SURVEY.md
section 8(f) "next" rows) are not executed: their outputs are stood in by fixed synthetic tensors of the right
shape, so every hot-path plugin runs on full-size, data-dependent inputs (voxel counts, set indices and masks
come from the real voxeliser / partition of the frame's cloud).

``ffn`` widens the frame by SURVEY.md 8(f) #4: the two FFN linears of every encoder layer
(fullyConnected_gelu_fullyConnected, src/dsvt-ai-trt.cpp:494-529) run on the FP32-accurate tensor-core linear
kernel (dsvt_linear_rows_launch), so a DSVT block becomes a real data flow from its input rows to its output rows:
    "graph": FC 192->384 -> GeluPlugin -> FC 384->192        (the reference graph's three nodes)
    "fused": FC 192->384 with the GELU in its epilogue -> FC 384->192 in split-K form with the residual add behind the FFN
             in its epilogue (one pass less over the 384-wide rows, one launch instead of two accumulating ones)
    "epilogue": the LayerNorms move into the GEMM epilogues as well -- norm1(attention + x) into the attention's
             out-projection (dsvt_set_attention_fused_norm_launch), the two / three LayerNorms behind the FFN into the
             second FFN linear (dsvt_linear_rows_norm_launch, K = 384 in one pass): 5 kernels per encoder layer, the
             attention output and the FFN output never reach memory
    "layer": "kernel" with the attention's out-projection and norm1 folded in front of the FFN kernel
             (dsvt_attention_tail_ffn_launch): 3 kernels per encoder layer, neither the attention output nor src's FFN read
             goes through memory
    "kernel": "epilogue" with the whole FFN (both linears, the GELU, the norms) as ONE kernel (dsvt_ffn_fused_launch): the
             384-wide hidden rows stay in tensor memory; 4 kernels per encoder layer

The linear layers stand for TensorRT FullyConnected layers, which have no zero-tail contract (the engine computes all
max_pillars rows; rows beyond the valid count hold bias-only values there and are never read by a plugin): they are launched
with zero_tails = 0 and leave those rows untouched.  Every plugin output keeps the reference's zero tails.

``backbone=True`` (with ``ffn`` on) runs the remaining TensorRT-native layers of the 3-D backbone as well, so that the frame
is ONE data flow from the raw points to the BEV map (random-init weights, BatchNorm folded):
    PFN layer 0  Linear(10->96)+BN+ReLU on the decorated points           (src/dsvt-ai-trt.cpp:577, dsvt_small_linear_launch)
    PFN layer 1  Linear(192->192)+BN+ReLU on [points | per-pillar max]    (:583-587, dsvt_linear_rows_concat_launch)
    8 position-embedding MLPs  Linear(2->192)+BN+ReLU -> Linear(192->192) (:603-637) on the in-window coordinates
"""
import ctypes

import numpy as np
import torch

from . import capi


class FrameWeights:
    """Weights of the 3-D backbone: 8 attention layers, 28 LayerNorms, FFN / PFN / position-embedding linears.
    ``FrameWeights(cfg, seed)`` draws random-init values (the bench: no checkpoint is reachable offline);
    ``FrameWeights.from_wts(cfg, tensors)`` takes the reference's trained tensors (dsvt.wts names, oracle/wts.py)."""

    def __init__(self, cfg, seed=0, device="cuda", _host=None):
        rng = np.random.default_rng(seed)
        C = cfg.channel_num
        host = _host if _host is not None else self._random_host(cfg, rng)
        self.attn_host = host["attn"]               # the host arrays the device images are made from (for the tests)
        self.attn = [capi.AttentionWeights(*a, C, cfg.num_heads) for a in self.attn_host]
        # stand-ins for the two PFN layer outputs (src/dsvt-ai-trt.cpp:577-590), shared by all frame slots (read-only)
        g = torch.Generator(device="cpu").manual_seed(seed + 1)
        self.pfn_out = [torch.randn(cfg.max_points_num_voxel_filter, f, generator=g).to(device) for f in cfg.pfn_channels]
        self.gamma = torch.from_numpy(host["gamma"]).to(device)
        self.beta = torch.from_numpy(host["beta"]).to(device)
        # FFN linears (fullyConnected_gelu_fullyConnected, src/dsvt-ai-trt.cpp:494-529): created on first use
        self._ffn_host = host["ffn"]
        self._ffn = None
        # VFE (PFN layers) and position-embedding MLPs (src/dsvt-ai-trt.cpp:577-637): BatchNorm1d folded to (scale, shift)
        self.vfe_host = host["vfe"]
        self.pos_host = host["pos"]
        self._glue = None

    @staticmethod
    def _random_host(cfg, rng):
        C = cfg.channel_num
        attn = [((rng.standard_normal((3 * C, C)) * 0.06).astype(np.float32),
                 (rng.standard_normal(3 * C) * 0.02).astype(np.float32),
                 (rng.standard_normal((C, C)) * 0.06).astype(np.float32),
                 (rng.standard_normal(C) * 0.02).astype(np.float32)) for _ in range(cfg.num_blocks * 2)]
        n_ln = cfg.num_blocks * 7
        gamma = (1.0 + 0.1 * rng.standard_normal((n_ln, C))).astype(np.float32)
        beta = (0.1 * rng.standard_normal((n_ln, C))).astype(np.float32)
        ffn = [((rng.standard_normal((cfg.ffn_channel_num, C)) * 0.07).astype(np.float32),
                (rng.standard_normal(cfg.ffn_channel_num) * 0.02).astype(np.float32),
                (rng.standard_normal((C, cfg.ffn_channel_num)) * 0.05).astype(np.float32),
                (rng.standard_normal(C) * 0.02).astype(np.float32)) for _ in range(cfg.num_blocks * 2)]

        def bn(n):
            g_, var = 1.0 + 0.1 * rng.standard_normal(n), rng.uniform(0.5, 1.5, n)
            mean, b_ = 0.1 * rng.standard_normal(n), 0.1 * rng.standard_normal(n)
            scale = 0.5 * g_ / np.sqrt(var + 1e-5)
            return scale.astype(np.float32), (b_ - mean * scale).astype(np.float32)
        F0, F1 = cfg.pfn_channels
        vfe = {"pfn0": ((rng.standard_normal((F0, cfg.feature_num)) * 0.05).astype(np.float32),) + bn(F0),
               "pfn1": ((rng.standard_normal((F1, 2 * F0)) * 0.07).astype(np.float32),) + bn(F1)}
        pos = [[((rng.standard_normal((C, 2)) * 0.3).astype(np.float32),) + bn(C) +
                ((rng.standard_normal((C, C)) * 0.07).astype(np.float32),
                 (rng.standard_normal(C) * 0.02).astype(np.float32)) for _ in range(2)]
               for _ in range(cfg.num_blocks)]
        return {"attn": attn, "gamma": gamma, "beta": beta, "ffn": ffn, "vfe": vfe, "pos": pos}

    @classmethod
    def from_wts(cls, cfg, t, seed=0, device="cuda"):
        """``t``: {name: float32 array} as oracle/wts.read_wts returns it (in_proj split into .query/.key/.value like the
        reference's loadWeights_new, include/helper.h:367-433).  The wiring follows src/dsvt-ai-trt.cpp:577-1128:
        BatchNorm1d folded with eps 1e-5 (:284, :477; add_batchNorm1d_relu :99-147), the first position-embedding linear's
        bias folded into the BatchNorm shift, LayerNorms in graph order norm1, norm2, norm per encoder + residual_norm."""
        C, F = cfg.channel_num, cfg.ffn_channel_num

        def bn(prefix, bias=None, eps=1e-5):
            g_, b_ = t[prefix + ".weight"].astype(np.float64), t[prefix + ".bias"].astype(np.float64)
            mean, var = t[prefix + ".running_mean"].astype(np.float64), t[prefix + ".running_var"].astype(np.float64)
            scale = g_ / np.sqrt(var + eps)
            shift = b_ - mean * scale
            if bias is not None:                     # y = (W x + bias) * scale + shift
                shift = shift + bias.astype(np.float64) * scale
            return scale.astype(np.float32), shift.astype(np.float32)

        attn, ffn, gamma, beta, pos = [], [], [], [], []
        for blk in range(cfg.num_blocks):
            row = []
            for enc in (0, 1):
                p = f"module.backbone_3d.stage_0.{blk}.encoder_list.{enc}"
                a = p + ".win_attn.self_attn"
                w_in = np.concatenate([t[f"{a}.in_proj_weight.{k}"] for k in ("query", "key", "value")]).reshape(3 * C, C)
                b_in = np.concatenate([t[f"{a}.in_proj_bias.{k}"] for k in ("query", "key", "value")])
                attn.append((w_in, b_in, t[a + ".out_proj.weight"].reshape(C, C), t[a + ".out_proj.bias"]))
                ffn.append((t[p + ".win_attn.linear1.weight"].reshape(F, C), t[p + ".win_attn.linear1.bias"],
                            t[p + ".win_attn.linear2.weight"].reshape(C, F), t[p + ".win_attn.linear2.bias"]))
                for n in (".win_attn.norm1", ".win_attn.norm2", ".norm"):
                    gamma.append(t[p + n + ".weight"]); beta.append(t[p + n + ".bias"])
                e = f"module.backbone_3d.input_layer.posembed_layers.0.{blk}.{enc}.position_embedding_head"
                row.append((t[e + ".0.weight"].reshape(C, 2),) + bn(e + ".1", bias=t[e + ".0.bias"]) +
                           (t[e + ".3.weight"].reshape(C, C), t[e + ".3.bias"]))
            pos.append(row)
            r = f"module.backbone_3d.residual_norm_stage_0.{blk}"
            gamma.append(t[r + ".weight"]); beta.append(t[r + ".bias"])
        F0, F1 = cfg.pfn_channels
        vfe = {"pfn0": (t["module.vfe.pfn_layers.0.linear.weight"].reshape(F0, cfg.feature_num),) + bn("module.vfe.pfn_layers.0.norm"),
               "pfn1": (t["module.vfe.pfn_layers.1.linear.weight"].reshape(F1, 2 * F0),) + bn("module.vfe.pfn_layers.1.norm")}
        host = {"attn": [tuple(np.ascontiguousarray(x, dtype=np.float32) for x in a) for a in attn],
                "ffn": [tuple(np.ascontiguousarray(x, dtype=np.float32) for x in a) for a in ffn],
                "gamma": np.stack(gamma).astype(np.float32), "beta": np.stack(beta).astype(np.float32),
                "vfe": vfe, "pos": pos}
        return cls(cfg, seed=seed, device=device, _host=host)

    @property
    def glue(self):
        """Device weights of the VFE / position-embedding layers: pfn0, pfn1, pos[blk][enc] = (first, second)."""
        if self._glue is None:
            w0, s0, t0 = self.vfe_host["pfn0"]
            w1, s1, t1 = self.vfe_host["pfn1"]
            self._glue = {
                "pfn0": capi.SmallLinear(w0, s0, t0),
                # Linear (no bias) + BatchNorm folded into the GEMM's weights and bias, ReLU in its epilogue
                "pfn1": capi.Linear(w1 * s1[:, None], t1, precision=capi.DSVT_ATTN_FP32_TC),
                "pos": [[(capi.SmallLinear(a, sc, sh), capi.Linear(b2, bias2, precision=capi.DSVT_ATTN_FP32_TC))
                         for a, sc, sh, b2, bias2 in row] for row in self.pos_host]}
        return self._glue

    @property
    def ffn(self):
        """[(Linear 192->384, Linear 384->192)] per encoder layer, FP32-accurate tensor-core weights images."""
        if self._ffn is None:
            self._ffn = [(capi.Linear(w1, b1, precision=capi.DSVT_ATTN_FP32_TC),
                          capi.Linear(w2, b2, precision=capi.DSVT_ATTN_FP32_TC)) for w1, b1, w2, b2 in self._ffn_host]
        return self._ffn


class HotPathFrame:
    """Buffers + launch sequence for one frame slot (one CUDA stream owns one slot)."""

    def __init__(self, cfg, weights, precision=capi.DSVT_ATTN_FP32, seed=0, device="cuda", fuse_ln=True, share_plans=True,
                 ffn="off", skip=(), zero_tails=1, backbone=False, head=False):
        assert ffn in ("off", "graph", "fused", "epilogue", "kernel", "layer")
        if ffn == "layer" and precision != capi.DSVT_ATTN_FP32_TC:
            ffn = "kernel"          # the layer-tail kernel continues the FP32_TC pipeline's workspace; other precisions keep 4 kernels
        assert not backbone or ffn != "off", "backbone=True runs every layer: it needs the FFN linears on"
        self.backbone = backbone
        # diagnostic only (tools/ablate.py): plugin groups left out of the launch sequence to measure their marginal
        # cost with several frames in flight -- {"vox", "smax", "part", "plan", "attn", "ln", "gelu", "m2b", "fbox"} and, for the
        # backbone3d frame, {"pfn", "pos", "ln1", "ffn1", "ffn2", "lnc"}
        self.skip = frozenset(skip)
        # 1 = the reference's contract (every output zero beyond its valid count, as its per-enqueue memsets leave it);
        # 0 = rows beyond the counts are left untouched (no consumer in the graph reads them) -- bench "relaxed_tails" leg only
        self.zero_tails = int(zero_tails)
        self.cfg, self.w, self.precision, self.fuse_ln, self.ffn = cfg, weights, precision, fuse_ln, ffn
        # GEMM-pipeline attention: one plan per (window partition, axis), shared by the two layers that use it
        self.share_plans = share_plans and precision in (capi.DSVT_ATTN_FP32_TC, capi.DSVT_ATTN_FP16_GEMM)
        self.plans = {}
        g = torch.Generator(device="cpu").manual_seed(seed)
        mp, C, F = cfg.max_pillars_num, cfg.channel_num, cfg.ffn_channel_num
        self.points = torch.zeros(1, cfg.max_points_num, 4, dtype=torch.float32, device=device)
        self.points_size = torch.zeros(1, dtype=torch.int32, device=device)
        self.vox = capi.Points2Features(cfg, device=device, zero_tails=self.zero_tails)
        self.wp = [capi.WindowPartition(cfg, i, device=device, zero_tails=self.zero_tails) for i in (0, 1)]
        self.gs = [capi.GetSet(cfg, i, device=device, zero_tails=self.zero_tails) for i in (0, 1)]
        # stand-ins for the outputs of TensorRT-native glue layers (only the ones this frame kind does not compute)
        if not backbone:
            self.x0 = torch.randn(mp, C, generator=g).to(device)                       # VFE / PFN output
            self.pos = [[torch.randn(mp, C, generator=g).mul_(0.5).to(device) for _ in range(2)]
                        for _ in range(cfg.num_blocks)]                                 # 8 pos-embed MLP outputs
        if ffn == "off":
            self.ffn_hidden = torch.randn(mp, F, generator=g).to(device)               # FFN linear 192->384 output
            self.ffn_out = torch.randn(mp, C, generator=g).mul_(0.5).to(device)        # FFN linear 384->192 output
        cand = __import__("importlib").import_module(__package__ + ".synth").head_candidates(cfg.max_top_k, seed)
        self.cand = [torch.from_numpy(c).to(device)[None] for c in cand]           # CenterHead top-K outputs
        # activations
        self.attn_out = torch.empty(mp, C, device=device)
        ws = capi.set_attention_workspace_bytes(1, cfg.max_win_num, cfg.voxel_num_set, C, cfg.num_heads, mp, precision)
        self.attn_ws = torch.empty(ws, dtype=torch.uint8, device=device) if ws else None   # qkv + o of the GEMM pipeline
        self.src = torch.empty(mp, C, device=device)
        self.src_b = torch.empty(mp, C, device=device)
        self.gelu_out = torch.empty(mp, F, device=device)
        if ffn != "off":
            self.ffn_h = torch.empty(mp, F, device=device)      # FC 192->384 output (graph form only)
            self.ffn_o = torch.empty(mp, C, device=device)      # FC 384->192 output
            self.ffn_parts = torch.empty(F // C, mp, C, device=device)    # ... as split-K partial sums (fused form)
        vfe_one = backbone and ffn in ("kernel", "layer")     # fused VFE kernel: the per-point tensors of the pillar feature net do not exist
        self.pfn0_out = self.pfn1_out = None
        if backbone:
            Pm = cfg.max_points_num_voxel_filter
            if not vfe_one:
                self.pfn0_out = torch.empty(Pm, cfg.pfn_channels[0], device=device)
                self.pfn1_out = torch.empty(Pm, cfg.pfn_channels[1], device=device)
            self.pos_hidden = None if vfe_one else torch.empty(mp, C, device=device)
            self.pos_out = None if ffn == "layer" else [[torch.empty(mp, C, device=device) for _ in range(2)] for _ in range(cfg.num_blocks)]
            if ffn == "layer":
                # position embedding as a TABLE over the cells of a window (its MLP's input is a function of (cx, cy) only,
                # windowPartition.cu:358-359): evaluated per frame on the list of cells, looked up by the QKV kernel
                ncell = max(wx * wy for wx, wy, _ in cfg.win_shapes)
                self.pos_cells = []
                for wx, wy, _ in cfg.win_shapes:
                    c = np.zeros((ncell, 2), np.float32)
                    cy, cx = np.divmod(np.arange(wx * wy), wx)
                    c[: wx * wy, 0] = cx.astype(np.float32) - np.float32(wx) / 2
                    c[: wx * wy, 1] = cy.astype(np.float32) - np.float32(wy) / 2
                    self.pos_cells.append(torch.from_numpy(c).to(device))
                self.pos_rows = torch.tensor([ncell], dtype=torch.int32, device=device)
                self.pos_tab = [[torch.empty(ncell, C, device=device) for _ in range(2)] for _ in range(cfg.num_blocks)]
        self.x_a = torch.empty(mp, C, device=device)
        self.x_b = torch.empty(mp, C, device=device)
        self.blk_out = [torch.empty(mp, C, device=device) for _ in range(2)]
        # VFE glue plugins (TorchScatterMaxPlugin x2) and the BEV map (Map2BevPlugin)
        self.max_point = None if vfe_one else [torch.empty(cfg.max_points_num_voxel_filter, f, device=device) for f in cfg.pfn_channels]
        self.max_voxel = [None if (vfe_one and i == 0) else torch.empty(mp, f, device=device) for i, f in enumerate(cfg.pfn_channels)]
        self.bev = torch.empty(cfg.grid_y, cfg.grid_x, C, device=device)
        self.boxes = torch.empty(1, cfg.max_top_k, 9, device=device)
        self.valid = torch.empty(1, dtype=torch.int32, device=device)
        # head=True: the post-process graph behind the CenterHead (sigmoid / TopK / gathers, src/dsvt-ai-trt.cpp:1471-1691) on
        # synthetic head maps (the 2-D backbone + CenterHead convolutions are not executed), FilterBoxByScorePlugin on ITS
        # outputs instead of on synthetic candidates, and the rotated NMS the reference runs on the host (helper.h:257-283)
        self.head = head
        if head:
            nc = 10
            gh = torch.Generator(device="cpu").manual_seed(seed + 77)
            r = lambda c, mul=1.0, add=0.0: torch.randn(1, c, cfg.grid_y, cfg.grid_x, generator=gh).mul_(mul).add_(add).to(device)
            # heat-map logits with ~250 cells above the 0.3 score threshold, centre offsets in [0,1), log sizes, (cos, sin)
            self.head_maps = (r(nc, 1.0, -4.6), torch.rand(1, 2, cfg.grid_y, cfg.grid_x, generator=gh).to(device),
                              r(1, 0.8, -1.0), r(3, 0.4, 0.5), r(2))
            # head="conv": the head maps come from the cuDNN stand-in of the 2-D backbone + CenterHead convolutions run on the
            # frame's own BEV map (conv_standin.py: library code with random weights, labelled; bench "whole_pipeline" leg)
            self.conv = None
            if head == "conv":
                cs = __import__("importlib").import_module(__package__ + ".conv_standin")
                self.conv = cs.BevHeadStandIn(cfg.grid_y, cfg.grid_x, seed=seed + 78, device=device)
            self.topk = capi.CenterHeadTopK(nc, cfg.grid_y, cfg.grid_x, cfg.max_top_k, device=device)
            self.nms = capi.RotatedNms(cfg.max_top_k, 0.01, device=device, zero_tails=self.zero_tails)      # NMS_THRESH, params.h:334
        self.launches_per_frame = None
        self.vfe_ws = None

    def attn_pos(self, blk, enc):
        """(pos, pos_table) arguments of capi.set_attention_fused for encoder layer (blk, enc) in this frame kind."""
        if self.ffn == "layer":
            return None, (self.pos_tab[blk][enc], self.wp[enc].coors_in_win_2d[0], self.cfg.win_shapes[enc][0])
        return self.pos_out[blk][enc], None

    def calibrate_head(self, n_above=250):
        """head="conv" only, random weights: shift the heat-map bias so that n_above cells of THIS frame's map score above
        SCORE_THRESH (a trained head's map is that sparse); one frame run + sync, outside any timed region."""
        self.run()
        torch.cuda.synchronize()
        m = self.conv(self.bev)["hm"].flatten()
        kth = torch.topk(m, n_above).values[-1].item()
        logit = float(np.log(self.cfg.score_threshold / (1.0 - self.cfg.score_threshold)))
        self.conv.heads["hm"][1][1].add_(logit - kth)

    def load_points(self, pts_np):
        n = min(len(pts_np), self.cfg.max_points_num)
        self.points[0, :n].copy_(torch.from_numpy(np.ascontiguousarray(pts_np[:n])))
        self.points_size.fill_(n)

    def run(self):
        """Enqueue the frame's plugin invocations on the current stream."""
        cfg, w = self.cfg, self.w
        before = capi.launch_count()
        skip, zt = self.skip, self.zero_tails
        vox = self.vox if "vox" in skip else self.vox(self.points, self.points_size)
        V = vox.pillar_num
        vfe_one = self.backbone and self.ffn in ("kernel", "layer")      # PFN 0 + scatter-max + concat + PFN 1 + scatter-max in one kernel
        if vfe_one and not ("smax" in skip and "pfn" in skip):
            g = w.glue
            if self.vfe_ws is None:
                self.vfe_ws = torch.empty(int(capi._lib().dsvt_vfe_fused_workspace_size(
                    ctypes.c_int32(vox.point_features.shape[-2]), ctypes.c_int32(vox.point_index_in_voxel.shape[-1]))) + 16,
                    dtype=torch.uint8, device=self.points.device)
            capi.vfe_fused(g["pfn0"], g["pfn1"], vox.point_features[0], vox.point_index_in_voxel[0], V, vox.point_num,
                           out=self.max_voxel[-1], workspace=self.vfe_ws, zero_tails=zt)
        for k in range(0 if ("smax" in skip or vfe_one) else len(cfg.pfn_channels)):             # :580-590 (the voxeliser's row count lets it skip the full clear)
            pfn_out = self.w.pfn_out[k]
            if self.backbone:                                                       # the PFN layer in front of the scatter-max
                g = w.glue
                if "pfn" in skip:
                    pfn_out = self.pfn0_out if k == 0 else self.pfn1_out
                elif k == 0:
                    pfn_out = g["pfn0"](vox.point_features[0], vox.point_num, activation=2, out=self.pfn0_out, zero_tails=0)
                else:
                    pfn_out = g["pfn1"].rows_concat(self.pfn0_out, self.max_point[0], vox.point_num, activation=2,
                                                    out=self.pfn1_out, zero_tails=0)
            capi.torch_scatter_max(pfn_out, vox.point_index_in_voxel[0], vox.point_num_in_voxel[0], V,
                                   vox.point_num, max_point=self.max_point[k], max_voxel=self.max_voxel[k], zero_tails=zt)
        for i in (() if "part" in skip else (0, 1)):
            self.wp[i](vox.coords, V)
            self.gs[i](self.wp[i].global_index, self.wp[i].coors_in_win, self.wp[i].voxel_num_in_win,
                       self.wp[i].win_num)
        if self.share_plans and not ("plan" in skip and self.plans):
            for part in (0, 1):
                for axis in (0, 1):
                    gs = self.gs[part]
                    self.plans[(part, axis)] = capi.set_attention_plan(
                        gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num, axis, cfg.max_pillars_num,
                        cfg.num_heads, cfg.channel_num, out=self.plans.get((part, axis)))
        x, ln = None, 0
        if not self.backbone:
            x, pos = self.x0, self.pos
        else:
            x = self.max_voxel[-1]                         # VFE output: per-pillar max of PFN layer 1 (:589, output 1)
            pos = self.pos_out
            if self.ffn == "layer" and "pos" not in skip:    # the MLPs on the cells of a window: tables of <= 24 x 24 rows
                pairs = [(blk, enc) for blk in range(cfg.num_blocks) for enc in (0, 1)]
                for i0 in range(0, len(pairs), 8):
                    grp = pairs[i0:i0 + 8]
                    capi.pos_embed_mlp_batch([w.glue["pos"][b_][e_][0] for b_, e_ in grp], [w.glue["pos"][b_][e_][1] for b_, e_ in grp],
                                             [self.pos_cells[e_] for b_, e_ in grp], self.pos_rows,
                                             [self.pos_tab[b_][e_] for b_, e_ in grp], zero_tails=0)
            if self.ffn == "kernel" and "pos" not in skip:   # all MLPs of the frame in one launch (they depend on the coordinates only)
                pairs = [(blk, enc) for blk in range(cfg.num_blocks) for enc in (0, 1)]
                for i0 in range(0, len(pairs), 8):
                    grp = pairs[i0:i0 + 8]
                    capi.pos_embed_mlp_batch([w.glue["pos"][b_][e_][0] for b_, e_ in grp], [w.glue["pos"][b_][e_][1] for b_, e_ in grp],
                                             [self.wp[e_].coors_in_win_x_y[0] for b_, e_ in grp], V,
                                             [self.pos_out[b_][e_] for b_, e_ in grp], zero_tails=0)
            for blk in range(0 if ("pos" in skip or self.ffn in ("kernel", "layer")) else cfg.num_blocks):   # pos_embed[blk][i] from the shift-i window coordinates (:603-637)
                for enc in (0, 1):
                    first, second = w.glue["pos"][blk][enc]
                    if self.ffn in ("epilogue", "kernel"):  # both layers in one kernel: the hidden rows never reach memory
                        capi.pos_embed_mlp(first, second, self.wp[enc].coors_in_win_x_y[0], V, out=self.pos_out[blk][enc],
                                           zero_tails=0)
                        continue
                    first(self.wp[enc].coors_in_win_x_y[0], V, activation=2, out=self.pos_hidden, zero_tails=0)
                    second.rows(self.pos_hidden, V, out=self.pos_out[blk][enc], zero_tails=0)
        for blk in range(cfg.num_blocks):
            gs = self.gs[blk % 2]                      # blocks 0,2: 12x12 windows; 1,3: 24x24 shifted (:654-:1018)
            x_in = x
            for enc in (0, 1):
                if self.ffn == "layer":
                    # three kernels per encoder layer: QKV projection, per-set core, then out-projection + norm1 + FFN + norms
                    aw, plan = w.attn[blk * 2 + enc], self.plans.get((blk % 2, enc))
                    st3 = 3 - sum(bit for g_, bit in (("attn_qkv", 1), ("attn_core", 2)) if g_ in skip)
                    if "attn" not in skip and st3:
                        capi.set_attention_fused(aw, x, None, gs.global_index_in_set[0], gs.mask_expand_0[0], gs.set_num, V,
                                                 axis=enc, out=self.src, precision=self.precision, workspace=self.attn_ws,
                                                 plan=plan, zero_tails=zt, stages=st3,
                                                 pos_table=(self.pos_tab[blk][enc], self.wp[enc].coors_in_win_2d[0], cfg.win_shapes[enc][0]))
                    fc1, fc2 = w.ffn[blk * 2 + enc]
                    norm1 = (w.gamma[ln], w.beta[ln], cfg.layer_norm_eps)
                    stages = [(self.src, w.gamma[ln + 1], w.beta[ln + 1]), (x, w.gamma[ln + 2], w.beta[ln + 2])]
                    ln += 3
                    if enc == 1:
                        stages.append((x_in, w.gamma[ln], w.beta[ln])); ln += 1
                    nxt = self.x_a if enc == 0 else self.blk_out[blk % 2]
                    if not ({"attn", "attn_out", "ffn2"} & skip):
                        capi.attention_tail_ffn(aw, fc1, fc2, x, gs.global_index_in_set[0], V, enc, plan, self.attn_ws, norm1, stages,
                                                cfg.layer_norm_eps, src=self.src, out=nxt, zero_tails=zt)
                    x = nxt
                    continue
                epi = self.ffn in ("epilogue", "kernel")
                if "attn" not in skip:
                    stages = 7 - sum(bit for g_, bit in (("attn_qkv", 1), ("attn_core", 2), ("attn_out", 4)) if g_ in skip)
                    capi.set_attention_fused(w.attn[blk * 2 + enc], x, pos[blk][enc], gs.global_index_in_set[0],
                                             gs.mask_expand_0[0], gs.set_num, V, axis=enc,
                                             out=self.src if epi else self.attn_out,
                                             precision=self.precision, workspace=self.attn_ws,
                                             plan=self.plans.get((blk % 2, enc)), zero_tails=zt,
                                             norm=(x, w.gamma[ln], w.beta[ln], cfg.layer_norm_eps) if epi else None,
                                             stages=stages)
                if "ln" in skip:
                    ln += 3 if enc == 0 else 4
                    if "gelu" not in skip:
                        capi.gelu(self.ffn_hidden, V, out=self.gelu_out, zero_tails=zt)
                    continue
                if "ln1" not in skip and not epi:
                    capi.layer_norm(self.attn_out, V, w.gamma[ln], w.beta[ln], cfg.layer_norm_eps, residual=x,
                                    out=self.src, zero_tails=zt)                           # norm1(y + x)   :669-676
                ln += 1
                ffn_out = None
                if self.ffn == "off":
                    ffn_out = self.ffn_out
                    if "gelu" not in skip:
                        capi.gelu(self.ffn_hidden, V, out=self.gelu_out, zero_tails=zt)    # :519 (inside the FFN)
                else:
                    fc1, fc2 = w.ffn[blk * 2 + enc]
                    if self.ffn == "graph":
                        fc1.rows(self.src, V, out=self.ffn_h, zero_tails=0)               # :513  FC 192->384
                        capi.gelu(self.ffn_h, V, out=self.gelu_out, zero_tails=zt)         # :519  GeluPlugin
                    elif "ffn1" not in skip and self.ffn != "kernel":
                        fc1.rows(self.src, V, activation=1, out=self.gelu_out, zero_tails=0)   # FC + GELU epilogue
                    if epi:
                        # FC 384->192 (one pass over K) + norm2(src + src2) + norm(src + x) [+ the block's residual norm]
                        stages = [(self.src, w.gamma[ln], w.beta[ln]), (x, w.gamma[ln + 1], w.beta[ln + 1])]
                        ln += 2
                        if enc == 1:
                            stages.append((x_in, w.gamma[ln], w.beta[ln])); ln += 1
                        nxt = self.x_a if enc == 0 else self.blk_out[blk % 2]
                        if self.ffn == "kernel":
                            if "ffn2" not in skip:          # both linears, the GELU and the norms in one kernel
                                fc1.ffn_norm(fc2, self.src, V, stages, cfg.layer_norm_eps, out=nxt, zero_tails=zt)
                        elif "ffn2" not in skip:
                            fc2.rows_norm(self.gelu_out, V, stages, cfg.layer_norm_eps, out=nxt, zero_tails=zt)
                        x = nxt
                        continue
                    if self.ffn == "graph":
                        ffn_out = fc2.rows(self.gelu_out, V, out=self.ffn_o, zero_tails=0)   # :524  FC 384->192
                nxt = self.x_a if enc == 0 else self.x_b
                ln_in = self.src
                if self.ffn == "fused" and "ln" not in skip:
                    # FC 384->192 in split-K form: ONE launch, part 0 = first K block + bias + src (the residual add behind
                    # the FFN folded into the epilogue), part 1 = second K block; norm2 sums them
                    parts = self.ffn_parts if "ffn2" in skip else fc2.rows_splitk(self.gelu_out, V, add=self.src, out=self.ffn_parts)
                    ln_in, ffn_out = parts[0], parts[1]
                if not self.fuse_ln:
                    capi.layer_norm(ln_in, V, w.gamma[ln], w.beta[ln], cfg.layer_norm_eps, residual=ffn_out,
                                    out=self.src_b, zero_tails=zt); ln += 1                # norm2(src + src2) :685-690
                    capi.layer_norm(self.src_b, V, w.gamma[ln], w.beta[ln], cfg.layer_norm_eps, residual=x,
                                    out=nxt, zero_tails=zt); ln += 1                       # norm(src + x)  :691-697
                    if enc == 1:
                        capi.layer_norm(nxt, V, w.gamma[ln], w.beta[ln], cfg.layer_norm_eps, residual=x_in,
                                        out=self.blk_out[blk % 2], zero_tails=zt); ln += 1  # residual norm  :750-756
                else:
                    # the same LayerNorms as one chained launch (rows stay in registers between stages)
                    stages = [(ffn_out, w.gamma[ln], w.beta[ln]), (x, w.gamma[ln + 1], w.beta[ln + 1])]
                    ln += 2
                    if enc == 1:
                        stages.append((x_in, w.gamma[ln], w.beta[ln])); ln += 1
                    if "lnc" not in skip:
                        capi.layer_norm_chain(ln_in, V, stages, cfg.layer_norm_eps,
                                              out=nxt if enc == 0 else self.blk_out[blk % 2], zero_tails=zt)
                x = nxt
            x = self.blk_out[blk % 2]
        self.final = x
        if "m2b" not in skip:
            capi.map2bev(x, vox.coords[0], V, cfg.grid_x, cfg.grid_y, out=self.bev)      # :1128
        if self.head:
            if self.conv is not None:
                m = self.conv(self.bev)
                self.topk(m["hm"], m["center"], m["center_z"], m["dim"], m["rot"])
            else:
                self.topk(*self.head_maps)
            capi.filter_box(cfg, *self.topk.outputs, boxes=self.boxes, valid=self.valid, zero_tails=zt)
            self.nms(self.boxes, self.valid)
        elif "fbox" not in skip:
            capi.filter_box(cfg, *self.cand, boxes=self.boxes, valid=self.valid, zero_tails=zt)
        self.launches_per_frame = capi.launch_count() - before
        return self
