"""Build recipe for the oracle.

* ``build_oracle()``  -- gcc-compiles ``dsvt_oracle.c`` (the C restatement) into
  ``oracle/_build/libdsvt_oracle.so``.
* ``build_reference()`` -- when ``/root/reference`` is present (authoring container only),
  nvcc-compiles the reference's OWN plugin sources, unmodified and in place, against the in-repo
  TensorRT scaffold header, one shared object per plugin (they define clashing global helpers,
  SURVEY.md 2.1), each linked with ``dsvt-ai-trt_b200/csrc/plugins/plugin_c_api.cpp`` (the C harness of
  ``include/dsvt_b200_plugin_c.h``, which drives a plugin through its IPluginCreator / IPluginV2DynamicExt interface).  Outputs go only to ``oracle/_ref/``
  (git-ignored, shipped to the GPU box by gpurun).  No reference source is copied into the repo.
  The reference's own CMake build is NOT used (it needs TensorRT + Boost, both absent).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
STUB = os.path.join(ROOT, "dsvt-ai-trt_b200", "csrc", "trt_stub")
PLUGIN_DIR = os.path.join(ROOT, "dsvt-ai-trt_b200", "csrc", "plugins")

REF_PLUGINS = [
    # (source stem, plugin name registered by the reference)
    ("points2Features", "Points2FeaturesPlugin"),
    ("windowPartition", "WindowPartitionPlugin"),
    ("getSet", "GetSetPlugin"),
    ("getValueByIndex", "GetValueByIndexPlugin"),
    ("mapSetFeature2voxel", "MapSetFeature2VoxelPlugin"),
    ("layerNorm", "LayerNormPlugin"),
    ("gelu", "GeluPlugin"),
    ("filterBoxByScore", "FilterBoxByScorePlugin"),
    ("torchScatterMax", "TorchScatterMaxPlugin"),
    ("map2bev", "Map2BevPlugin"),
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build_oracle(verbose=False):
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(HERE, "dsvt_oracle.c")
    out = os.path.join(out_dir, "libdsvt_oracle.so")
    if _newer(out, [src]):
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
               "-fvisibility=hidden", "-Wall", "-o", out, src, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return out


def reference_available():
    return os.path.isdir(os.path.join(REF, "plugins", "src"))


def build_reference(verbose=False, variant=None):
    """Returns {plugin stem: path to .so}; only stems that built.

    variant=None    : the reference exactly as shipped (its own include/params.h; 50k points / 10k pillars / 800 sets)
    variant="waymo" : same unmodified sources, configured through oracle/ref_config_waymo/params.h placed first on
                      the include path (320k points / 40k pillars / 4096 sets) -> oracle/_ref/waymo/
    """
    out_dir = os.path.join(HERE, "_ref") if variant is None else os.path.join(HERE, "_ref", variant)
    cfg_inc = [] if variant is None else ["-I" + os.path.join(HERE, "ref_config_" + variant)]
    built = {}
    if not reference_available():
        # GPU box: use whatever was prebuilt and shipped
        for stem, _ in REF_PLUGINS:
            so = os.path.join(out_dir, f"libref_{stem}.so")
            if os.path.exists(so):
                built[stem] = so
        return built
    if shutil.which("nvcc") is None:
        return built
    os.makedirs(out_dir, exist_ok=True)
    harness = os.path.join(PLUGIN_DIR, "plugin_c_api.cpp")
    for stem, _ in REF_PLUGINS:
        src = os.path.join(REF, "plugins", "src", f"{stem}.cu")
        so = os.path.join(out_dir, f"libref_{stem}.so")
        deps = [src, harness, os.path.join(STUB, "NvInfer.h")]
        if variant is not None:
            deps.append(os.path.join(HERE, "ref_config_" + variant, "params.h"))
        if _newer(so, deps):
            cmd = ["nvcc", "-std=c++14", "-O2", "-w", "-gencode", "arch=compute_100a,code=sm_100a",
                   "-Xcompiler", "-fPIC", "-shared", *cfg_inc,
                   "-I" + STUB, "-I" + os.path.join(REF, "include"), "-I" + os.path.join(REF, "plugins", "include"),
                   "-I" + os.path.join(ROOT, "include"),
                   "-DDSVT_HARNESS_FOR_REFERENCE=1",
                   "-x", "cu", src, harness, "-o", so, "-lcudart", "-Xlinker", "-Bsymbolic"]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(f"[oracle/_ref] {stem}: build failed\n{r.stderr[-2000:]}\n")
                continue
        built[stem] = so
    return built


def build_reference_nms(verbose=False):
    """The reference's host-side rotated NMS (include/helper.h:92-283), compiled unmodified behind a 12-line C entry point
    (oracle/ref_nms_harness.cpp) -> oracle/_ref/libref_nms.so.  Returns the path or None."""
    so = os.path.join(HERE, "_ref", "libref_nms.so")
    if not reference_available():
        return so if os.path.exists(so) else None
    os.makedirs(os.path.dirname(so), exist_ok=True)
    harness = os.path.join(HERE, "ref_nms_harness.cpp")
    hdr = os.path.join(REF, "include", "helper.h")
    if _newer(so, [harness, hdr]):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        cmd = ["g++", "-std=c++14", "-O2", "-w", "-fPIC", "-shared", "-ffp-contract=off", "-I" + STUB, "-I" + os.path.join(REF, "include"),
               "-I" + cuda_inc, harness, "-o", so]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(f"[oracle/_ref] nms: build failed\n{r.stderr[-2000:]}\n")
            return None
    return so


if __name__ == "__main__":
    print(build_oracle(verbose=True))
    for variant in (None, "waymo"):
        for k, v in build_reference(verbose=True, variant=variant).items():
            print(variant, k, v)
    print(build_reference_nms(verbose=True))
