"""CPU-side checks of the drop-in boundary: the shared libraries load and export every symbol the
headers declare; creators are registered under the reference's names with the reference's fields;
serialised layouts match.  No kernel is launched here."""
import ctypes
import importlib
import os
import re
import struct

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsvt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    names = declared_functions("dsvt_b200.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libdsvt_b200.so does not export {n}"
    lib.dsvt_abi_version.restype = ctypes.c_int
    assert lib.dsvt_abi_version() == 1


def test_plugin_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_plugin_library()
    for n in declared_functions("dsvt_b200_plugin_c.h"):
        assert hasattr(lib, n), f"libdsvt_b200_plugins.so does not export {n}"


REFERENCE_FIELDS = {   # creator field lists of the reference (plugins/src/*.cu)
    "Points2FeaturesPlugin": ["max_points_num", "max_points_num_voxel_filter", "max_pillars_num", "point_feature_num",
                              "feature_num", "max_num_points_per_voxel", "point_cloud_range", "voxel_size", "grid_size"],
    "GetSetPlugin": ["max_win_num", "max_voxel_num_per_win", "voxel_num_set", "win_shape"],
    "GeluPlugin": ["max_pillars_num", "channel_num"],
    "LayerNormPlugin": ["max_pillars_num", "channel_num", "weights_size", "pes", "weights", "bias"],
    "FilterBoxByScorePlugin": ["max_top_k", "point_cloud_range", "voxel_size", "score_threshold"],
    "WindowPartitionPlugin": ["max_win_num", "max_voxel_num_per_win", "sparse_shape", "win_shape", "shift_list"],
    "GetValueByIndexPlugin": ["max_win_num", "voxel_num_set", "channel_num", "axis_id"],
    "MapSetFeature2VoxelPlugin": ["max_win_num", "voxel_num_set", "channel_num", "axis_id", "max_pillars_num"],
    "TorchScatterMaxPlugin": ["max_points_num", "max_pillars_num", "feature_num"],      # torchScatterMax.cu:388-390
    "Map2BevPlugin": ["max_pillars_num", "channel_num", "grid_size_x", "grid_size_y"],  # map2bev.cu:383-386
}


def test_creators_registered_with_reference_names_and_fields():
    plg = importlib.import_module("dsvt-ai-trt_b200.plugins")
    lib = plg.PluginLibrary()
    registered = lib.registered()
    for name, fields in REFERENCE_FIELDS.items():
        assert name in registered
        assert lib.field_names(name, "1") == fields
    with pytest.raises(KeyError):
        lib.field_names("NoSuchPlugin")
    with pytest.raises(KeyError):
        lib.field_names("GetSetPlugin", "2")


def test_serialisation_layouts_without_gpu():
    """Plugins that own no device memory can be created, serialised and re-created on a CPU-only box."""
    plg = importlib.import_module("dsvt-ai-trt_b200.plugins")
    lib = plg.PluginLibrary()
    p = plg.add_voxel_generator(lib, 50000, 30000, 10000, 4, 10, 48, -74.88, 74.88, -74.88, 74.88, -5.0, 3.0,
                                0.32, 0.32, 8.0, 468, 468, 1)
    assert p.serialize() == struct.pack("<6i9f3i", 50000, 30000, 10000, 4, 10, 48, -74.88, 74.88, -74.88, 74.88,
                                        -5.0, 3.0, 0.32, 0.32, 8.0, 468, 468, 1)
    g = plg.add_get_set_op(lib, 800, 576, 36, (12, 12, 1))
    blob = g.serialize()
    assert blob == struct.pack("<6i", 36, 800, 576, 12, 12, 1)
    g2 = lib.deserialize("GetSetPlugin", blob)
    assert g2.serialize() == blob and g2.nb_outputs == 5 and g2.type == "GetSetPlugin"
    f = plg.add_filter_box_by_score_op(lib, 500, -74.88, 74.88, -74.88, 74.88, -5.0, 3.0, 0.32, 0.32, 8.0, 0.3)
    assert f.serialize() == struct.pack("<i10f", 500, -74.88, 74.88, -74.88, 74.88, -5.0, 3.0, 0.32, 0.32, 8.0, 0.3)
    with pytest.raises(RuntimeError):
        lib.deserialize("GetSetPlugin", blob[:10])      # truncated plan data is rejected, not read out of bounds
    sm = plg.add_torch_scatter_max(lib, 30000, 10000, 96)
    assert sm.serialize() == struct.pack("<3i", 30000, 10000, 96) and sm.nb_outputs == 2    # torchScatterMax.cu:346-352
    mb = plg.add_map_2_bev_op(lib, 10000, 192, 468, 468)
    assert mb.serialize() == struct.pack("<4i", 10000, 192, 468, 468)                       # map2bev.cu:352-359
    assert lib.deserialize("Map2BevPlugin", mb.serialize()).serialize() == mb.serialize()
    # static output shapes (getOutputDimensions) and I/O formats (supportsFormatCombination)
    D = plg._Desc
    def desc(dims, dt):
        d = D(); d.nb_dims = len(dims)
        for i, v in enumerate(dims): d.dims[i] = v
        d.dtype = dt
        return d
    ins = [desc((1, 800, 576), 3), desc((1, 800, 576, 3), 3), desc((1, 800), 3), desc((1,), 3)]
    outs = g.output_descs(ins)
    assert [tuple(o.dims[: o.nb_dims]) for o in outs] == [(1, 2, 800, 36), (1, 2, 800, 36), (1,), (1, 800, 8, 36), (1, 800, 8, 36)]
    assert [o.dtype for o in outs] == [3, 0, 3, 0, 0]
    assert all(g.supports_format(i, ins + outs, 4) for i in range(9))
    bad = list(ins + outs); bad[5] = desc((1, 2, 800, 36), 3)
    assert not g.supports_format(5, bad, 4)
