"""CPU tests of the .wts reader (oracle/wts.py = loadWeights_new, reference include/helper.h:328-439) and of the
committed trained-weights fixture it produced (tools/make_wts_fixture.py)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import wts


def test_reader_round_trip_and_in_proj_split(tmp_path):
    rng = np.random.default_rng(0)
    t = {"module.global_step": np.array([12345.0], np.float32),
         "a.self_attn.in_proj_weight": rng.standard_normal(3 * 6 * 6).astype(np.float32),
         "a.self_attn.in_proj_bias": np.array([0, -0.0, 1e-38, 3.4e38, -1.5, 2.5, 7, 8, 9], np.float32),
         "a.norm.weight": rng.standard_normal(6).astype(np.float32)}
    path = tmp_path / "t.wts"
    wts.write_wts(path, t)
    text = open(path).read().split("\n")
    assert text[0] == "4" and text[1].startswith("module.global_step 1 4640e400")   # big-endian hex of the f32 bits
    plain = wts.read_wts(path, split_in_proj=False)
    assert plain.keys() == t.keys()
    for k in t:
        assert plain[k].dtype == np.float32 and np.array_equal(plain[k].view(np.uint32), t[k].view(np.uint32))
    split = wts.read_wts(path)
    assert "a.self_attn.in_proj_weight" not in split
    w = t["a.self_attn.in_proj_weight"]
    for i, part in enumerate(("query", "key", "value")):     # rows 0..C-1 query, C..2C-1 key, 2C..3C-1 value
        assert np.array_equal(split[f"a.self_attn.in_proj_weight.{part}"], w[i * 36:(i + 1) * 36])
        assert np.array_equal(split[f"a.self_attn.in_proj_bias.{part}"], t["a.self_attn.in_proj_bias"][i * 3:(i + 1) * 3])
    only = wts.read_wts(path, keep=lambda n: n.endswith("norm.weight"))
    assert list(only) == ["a.norm.weight"]


def test_reader_rejects_bad_files(tmp_path):
    p = tmp_path / "bad.wts"
    p.write_text("0\n")
    with pytest.raises(ValueError):
        wts.read_wts(p)
    p.write_text("1\nx 3 3f800000 3f800000 \n")
    with pytest.raises(ValueError):
        wts.read_wts(p)


def test_trained_fixture():
    t = dict(np.load(os.path.join(GOLDEN, "dsvt_backbone3d_wts.npz")))
    names = wts.backbone3d_names()
    assert len(t) == sum(3 if ".in_proj_" in n else 1 for n in names) == 226
    a = "module.backbone_3d.stage_0.3.encoder_list.1.win_attn.self_attn"
    assert t[a + ".in_proj_weight.key"].size == 192 * 192 and t[a + ".in_proj_bias.value"].size == 192
    assert t[a + ".out_proj.weight"].size == 36864 and t["module.vfe.pfn_layers.0.linear.weight"].size == 960
    assert sum(v.size for v in t.values()) == 2726976
    assert all(v.dtype == np.float32 and np.isfinite(v).all() for v in t.values())
    assert all(float(np.abs(v).max()) == 0.0 for k, v in t.items() if ".in_proj_bias." in k)   # zeros in the real file
    h = hashlib.sha256()
    for k in sorted(t):
        h.update(k.encode()); h.update(t[k].tobytes())
    assert h.hexdigest()[:16] == "d90def49169ee213"
    if os.path.exists("/root/reference/dsvt.wts"):           # build container only: the fixture equals the file
        sel = {a + ".in_proj_weight", "module.backbone_3d.residual_norm_stage_0.2.bias"}
        ref = wts.read_wts("/root/reference/dsvt.wts", keep=lambda n: n in sel)
        for k, v in ref.items():
            assert np.array_equal(v, t[k])
