"""Extracts the 3-D backbone's trained tensors from the reference's dsvt.wts into tests/golden/dsvt_backbone3d_wts.npz.

Run in the build container (the reference tree does not exist on the GPU box):
    python tools/make_wts_fixture.py [/root/reference/dsvt.wts]
The file is read with oracle/wts.py (the restatement of loadWeights_new, reference include/helper.h:328-439), i.e. the
in_proj tensors are stored split as .query / .key / .value exactly as the reference's weight map holds them.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wts  # noqa: E402


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/dsvt.wts"
    wanted = set(wts.backbone3d_names())
    t = wts.read_wts(src, split_in_proj=True, keep=lambda n: n in wanted)
    expect = sum(3 if ".in_proj_" in n else 1 for n in wanted)
    assert len(t) == expect, (len(t), expect)
    out = os.path.join(ROOT, "tests", "golden", "dsvt_backbone3d_wts.npz")
    np.savez_compressed(out, **t)
    h = hashlib.sha256()
    for k in sorted(t):
        h.update(k.encode()); h.update(t[k].tobytes())
    n = sum(v.size for v in t.values())
    print(f"{out}: {len(t)} tensors, {n} floats, {os.path.getsize(out) / 1e6:.1f} MB, sha256 {h.hexdigest()[:16]}")


if __name__ == "__main__":
    main()
