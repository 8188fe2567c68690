"""Generates tests/golden/* from the reference's own data (run in the authoring container only;
/root/reference does not exist on the GPU box).

* frame_000000.npz  -- the reference's sample cloud data/bin/000000.bin (md5 81ce6be5...), the input
                       its README run uses, stored losslessly (float32) so GPU tests can replay it.
* kat.json          -- known-answer table: per distinct reference frame the point /
                       in-range / pillar / kept-point / window / set / masked-slot counts.  5504 pillars and
                       454 sets also appear as shape comments in the reference source
                       (getValueByIndex.cu:176,184; dsvt-ai-trt.cpp:291).  The table is computed here with an
                       independent float32 numpy restatement (NOT with oracle/), so it pins the oracle.
* attention_case.npz -- small seeded set-attention case with the output of
                       torch.nn.functional.multi_head_attention_forward (the op the reference's
                       multHeadAttention() graph restates, src/dsvt-ai-trt.cpp:288-458).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def kat_for(points):
    f32 = np.float32
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    xmin, xmax, zmin, zmax, vs = f32(-74.88), f32(74.88), f32(-5.0), f32(3.0), f32(0.32)
    inr = (x >= xmin) & (x < xmax) & (y >= xmin) & (y < xmax) & (z >= zmin) & (z < zmax)
    ix = np.floor((x[inr] - xmin) / vs).astype(np.int64)     # float32 arithmetic (SURVEY A-1)
    iy = np.floor((y[inr] - xmin) / vs).astype(np.int64)
    cell = iy * 468 + ix
    cells, counts = np.unique(cell, return_counts=True)
    kept = int(np.minimum(counts, 48).sum())
    row = {"points": int(len(points)), "in_range": int(inr.sum()), "pillars": int(len(cells)),
           "overfull_pillars": int((counts > 48).sum()), "max_points_in_pillar": int(counts.max()), "kept_points": kept}
    cy, cx = cells // 468, cells % 468
    for tag, (w, s) in {"win12": (12, 0), "win24_shift6": (24, 6)}.items():
        nw = 468 // w + 1
        wid = ((cy + s) // w) * nw + (cx + s) // w
        wins, wc = np.unique(wid, return_counts=True)
        nsets = np.ceil(wc / 36).astype(np.int64)
        masked = 0
        for n, ns in zip(wc, nsets):     # slots whose rank repeats the previous slot's rank
            t = np.arange(ns * 36)
            r = (t * n // 36) // ns
            rep = (r[1:] == r[:-1]) & (t[1:] % 36 != 0)
            masked += int(rep.sum())
        row[tag] = {"windows": int(len(wins)), "max_voxels_per_window": int(wc.max()), "sets": int(nsets.sum()),
                    "masked_slots": masked, "slots": int(nsets.sum() * 36)}
    return row


def main():
    os.makedirs(OUT, exist_ok=True)
    table = {}
    for name in ("000000", "000003", "000004"):
        raw = open(os.path.join(REF, "data", "bin", name + ".bin"), "rb").read()
        pts = np.frombuffer(raw, dtype=np.float32).reshape(-1, 4)
        table[name] = {"md5": hashlib.md5(raw).hexdigest(), **kat_for(pts)}
        if name == "000000":
            np.savez_compressed(os.path.join(OUT, "frame_000000.npz"), points=pts)
    json.dump(table, open(os.path.join(OUT, "kat.json"), "w"), indent=1, sort_keys=True)

    import torch
    import torch.nn.functional as Fn
    g = torch.Generator().manual_seed(1234)
    sets, S, C, H = 4, 36, 192, 8
    q = torch.randn(sets, S, C, generator=g)
    k = q.clone()
    v = torch.randn(sets, S, C, generator=g)
    # weights are rounded to fp16-representable values so the fixture can store them in half the bytes
    w_in = (torch.randn(3 * C, C, generator=g) * 0.06).half().float()
    b_in = torch.randn(3 * C, generator=g) * 0.1
    w_out = (torch.randn(C, C, generator=g) * 0.06).half().float()
    b_out = torch.randn(C, generator=g) * 0.1
    masked = torch.zeros(sets, S, dtype=torch.bool)
    for s in range(sets):
        n_pad = int(torch.randint(0, 24, (1,), generator=g))
        idx = torch.randperm(S - 1, generator=g)[:n_pad] + 1      # slot 0 is never masked (SURVEY A-6 iv)
        masked[s, idx] = True
    mask = torch.where(masked, torch.tensor(-3.4028234663852886e38), torch.tensor(0.0))
    mask = mask[:, None, :].expand(sets, H, S).contiguous()
    out, _ = Fn.multi_head_attention_forward(
        q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1), C, H, w_in, b_in, None, None, False, 0.0,
        w_out, b_out, training=False, key_padding_mask=masked, need_weights=False)
    np.savez_compressed(os.path.join(OUT, "attention_case.npz"), q=q.numpy(), k=k.numpy(), v=v.numpy(),
                        mask=mask.numpy(), w_in=w_in.half().numpy(), b_in=b_in.numpy(), w_out=w_out.half().numpy(),
                        b_out=b_out.numpy(), out=out.transpose(0, 1).contiguous().numpy())
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    sys.exit(main())
