"""ctypes loaders for the two in-tree shared libraries.  No fallback of any kind."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DSVT_B200_LIBDIR: A/B builds of the same sources (tuning experiments); the default is the in-tree build
_LIB_DIR = os.environ.get("DSVT_B200_LIBDIR") or os.path.join(_HERE, "lib")


class LibraryMissing(RuntimeError):
    pass


def library_path():
    return os.path.join(_LIB_DIR, "libdsvt_b200.so")


def plugin_library_path():
    return os.path.join(_LIB_DIR, "libdsvt_b200_plugins.so")


_cache = {}


def _load(path):
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise LibraryMissing(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C dsvt-ai-trt_b200/csrc`).  There is no CPU fallback.")
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    _cache[path] = lib
    return lib


def load_library():
    """libdsvt_b200.so (kernels + C ABI)."""
    return _load(library_path())


def load_plugin_library():
    """libdsvt_b200_plugins.so (IPluginV2DynamicExt shells + C harness); depends on libdsvt_b200.so."""
    load_library()
    return _load(plugin_library_path())
