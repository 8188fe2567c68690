"""Reader of the reference's TensorRT weight file (dsvt.wts).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates ``loadWeights_new`` (reference include/helper.h:328-439; plain ``loadWeights`` :286-326 is the same
without the split): a text file, first token = number of blobs, then per blob ``<name> <decimal count> <count
hex words>``, every word the IEEE-754 bit pattern of one float32.  Blobs whose name contains ``.in_proj_`` are
split in three equal consecutive parts stored as ``<name>.query`` / ``.key`` / ``.value`` (helper.h:367-433: rows
0-191 = query, 192-383 = key, 384-575 = value of ``in_proj_weight`` / ``in_proj_bias``).  ``tools/gen_wts.py:88-101``
of the reference is the writer: ``struct.pack('>f', v).hex()`` per value.
"""
import numpy as np


def read_wts(path, split_in_proj=True, keep=None):
    """-> {name: float32 array}.  ``keep``: optional predicate on the blob name (skipped blobs are not decoded)."""
    out = {}
    with open(path, "r") as f:
        first = f.readline().split()
        count = int(first[0])
        if count <= 0:
            raise ValueError("Invalid weight map file.")          # helper.h:300 / :343 assert
        for _ in range(count):
            line = f.readline()
            if not line:
                raise ValueError("truncated weight file")
            name, size, rest = line.split(" ", 2)
            size = int(size)
            if keep is not None and not keep(name):
                continue
            words = rest.split()
            if len(words) != size:
                raise ValueError(f"{name}: {len(words)} words, header says {size}")
            vals = np.array([int(w, 16) for w in words], dtype=np.uint32).view(np.float32)
            if split_in_proj and ".in_proj_" in name:
                n = size // 3
                for i, part in enumerate(("query", "key", "value")):
                    out[f"{name}.{part}"] = vals[i * n:(i + 1) * n].copy()
            else:
                out[name] = vals
    return out


def write_wts(path, tensors):
    """Inverse of read_wts for un-split names (the reference's tools/gen_wts.py format) -- used by the tests."""
    with open(path, "w") as f:
        f.write(f"{len(tensors)}\n")
        for name, arr in tensors.items():
            a = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
            f.write(f"{name} {a.size} ")
            f.write(" ".join(f"{w:08x}" for w in a.view(np.uint32)))
            f.write(" \n")


def backbone3d_names(num_blocks=4):
    """Names (un-split, as in the file) of every tensor the 3-D backbone uses (src/dsvt-ai-trt.cpp:577-1128)."""
    names = []
    for i in (0, 1):
        names += [f"module.vfe.pfn_layers.{i}.linear.weight"]
        names += [f"module.vfe.pfn_layers.{i}.norm.{k}" for k in ("weight", "bias", "running_mean", "running_var")]
    for blk in range(num_blocks):
        for enc in (0, 1):
            p = f"module.backbone_3d.input_layer.posembed_layers.0.{blk}.{enc}.position_embedding_head"
            names += [f"{p}.0.weight", f"{p}.0.bias", f"{p}.3.weight", f"{p}.3.bias"]
            names += [f"{p}.1.{k}" for k in ("weight", "bias", "running_mean", "running_var")]
    for blk in range(num_blocks):
        for enc in (0, 1):
            p = f"module.backbone_3d.stage_0.{blk}.encoder_list.{enc}"
            names += [f"{p}.win_attn.self_attn.in_proj_weight", f"{p}.win_attn.self_attn.in_proj_bias",
                      f"{p}.win_attn.self_attn.out_proj.weight", f"{p}.win_attn.self_attn.out_proj.bias",
                      f"{p}.win_attn.linear1.weight", f"{p}.win_attn.linear1.bias",
                      f"{p}.win_attn.linear2.weight", f"{p}.win_attn.linear2.bias",
                      f"{p}.win_attn.norm1.weight", f"{p}.win_attn.norm1.bias",
                      f"{p}.win_attn.norm2.weight", f"{p}.win_attn.norm2.bias", f"{p}.norm.weight", f"{p}.norm.bias"]
        names += [f"module.backbone_3d.residual_norm_stage_0.{blk}.weight", f"module.backbone_3d.residual_norm_stage_0.{blk}.bias"]
    return names
