"""STAND-IN (library code, NOT part of the hot path and not a product kernel): the 2-D BEV backbone + CenterHead
convolutions of the reference (src/dsvt-ai-trt.cpp:1137-1468, shapes from include/params.h:86-322), which the reference
runs as TensorRT-native convolution layers.  Here they are cuDNN convolutions through PyTorch with random-init weights
(BatchNorm folded) in BF16 / channels-last, so that a whole-pipeline frame time -- raw points -> boxes after NMS -- can be
quoted next to the reference README's 0.7 s per frame.  SURVEY.md section 8 marks these layers out of scope; nothing here
counts as a kernel of this repo.

Structure (BaseBEVResBackbone + CenterHead of the DSVT-pillar nuScenes model):
  block 0: BasicBlock(192 -> 128, 1x1 downsample), BasicBlock(128)                          @ stride 1
  block 1: BasicBlock(128 -> 128, stride 2, 1x1 s2 downsample), 2 x BasicBlock(128)          @ stride 2
  block 2: BasicBlock(128 -> 256, stride 2, 1x1 s2 downsample), 2 x BasicBlock(256)          @ stride 4
  deblocks: ConvTranspose 1x1 s1 (128 -> 128), 2x2 s2 (128 -> 128), 4x4 s4 (256 -> 128), each + BN + ReLU; concat -> 384
  shared conv 3x3 (384 -> 64) + BN + ReLU; six heads: conv 3x3 (64 -> 64) + BN + ReLU, conv 3x3 (64 -> n)
  (center 2, center_z 1, dim 3, rot 2, iou 1, hm 10).
"""
import torch
import torch.nn.functional as F


class BevHeadStandIn:
    def __init__(self, grid_y, grid_x, seed=0, device="cuda", dtype=torch.bfloat16):
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.dtype, self.H, self.W = dtype, grid_y, grid_x

        def conv(cin, cout, k, transposed=False):
            shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
            w = torch.randn(*shape, generator=g) * (2.0 / (cin * k * k)) ** 0.5
            b = torch.randn(cout, generator=g) * 0.02
            return (w.to(device=device, dtype=dtype).contiguous(memory_format=torch.channels_last), b.to(device=device, dtype=dtype))

        def block(cin, cout, stride):
            return {"c1": conv(cin, cout, 3), "c2": conv(cout, cout, 3), "stride": stride,
                    "down": conv(cin, cout, 1) if (stride != 1 or cin != cout) else None}

        self.blocks = [[block(192, 128, 1), block(128, 128, 1)],
                       [block(128, 128, 2), block(128, 128, 1), block(128, 128, 1)],
                       [block(128, 256, 2), block(256, 256, 1), block(256, 256, 1)]]
        self.deblocks = [(conv(128, 128, 1, True), 1), (conv(128, 128, 2, True), 2), (conv(256, 128, 4, True), 4)]
        self.shared = conv(384, 64, 3)
        self.heads = {n: (conv(64, 64, 3), conv(64, c, 3)) for n, c in
                      (("center", 2), ("center_z", 1), ("dim", 3), ("rot", 2), ("iou", 1), ("hm", 10))}
        with torch.no_grad():      # sparse heat map like a trained head: a few hundred cells above the 0.3 score threshold
            self.heads["hm"][1][1].fill_(-6.0)
        self.flops = None

    @staticmethod
    def _basic(x, blk):
        y = F.relu(F.conv2d(x, *blk["c1"], stride=blk["stride"], padding=1))
        y = F.conv2d(y, *blk["c2"], padding=1)
        idn = x if blk["down"] is None else F.conv2d(x, *blk["down"], stride=blk["stride"])
        return F.relu(y + idn)

    def __call__(self, bev_hwc):
        """bev_hwc [H, W, 192] f32 (Map2BevPlugin's output) -> dict of FP32 NCHW head maps for dsvt_center_head_topk_launch."""
        x = bev_hwc.permute(2, 0, 1)[None].to(self.dtype)          # NCHW view of channels-last data: no transpose
        ups = []
        for blks, (dw, s) in zip(self.blocks, self.deblocks):
            for blk in blks:
                x = self._basic(x, blk)
            ups.append(F.relu(F.conv_transpose2d(x, *dw, stride=s))[..., : self.H, : self.W])
        y = F.relu(F.conv2d(torch.cat(ups, dim=1), *self.shared, padding=1))
        out = {}
        for name, (c0, c1) in self.heads.items():
            out[name] = F.conv2d(F.relu(F.conv2d(y, *c0, padding=1)), *c1, padding=1).float().contiguous()
        return out

    def gflop(self):
        """Dense multiply-add count of the stack x 2, in GFLOP (for the record next to the frame time)."""
        H, W, tot = self.H, self.W, 0.0
        dims = [(H, W), ((H + 1) // 2, (W + 1) // 2), ((H + 3) // 4, (W + 3) // 4)]
        for blks, (h, w) in zip(self.blocks, dims):
            for blk in blks:
                for key in ("c1", "c2", "down"):
                    if blk[key] is not None:
                        wt = blk[key][0]
                        tot += 2.0 * h * w * wt.shape[0] * wt.shape[1] * wt.shape[2] * wt.shape[3]
        for (dw, s), (h, w) in zip(self.deblocks, dims):
            tot += 2.0 * h * w * dw[0].shape[0] * dw[0].shape[1] * dw[0].shape[2] * dw[0].shape[3]
        tot += 2.0 * H * W * 64 * 384 * 9
        for c0, c1 in self.heads.values():
            tot += 2.0 * H * W * (64 * 64 * 9 + c1[0].shape[0] * 64 * 9)
        return tot * 1e-9
