"""dsvt-ai-trt_b200 -- B200-native (sm_100a) DSVT hot path behind the reference's plugin surface.

The directory name carries a hyphen (it mirrors the reference repo's name), so import it with
``importlib.import_module("dsvt-ai-trt_b200")`` or through the ``dsvt_b200`` shim at the repo root.

The compute path is the CUDA library ``lib/libdsvt_b200.so`` (C ABI: ``include/dsvt_b200.h``) and
the TensorRT plugin shells ``lib/libdsvt_b200_plugins.so`` (C harness: ``include/dsvt_b200_plugin_c.h``).
There is NO CPU fallback: loading fails loudly when the libraries are missing.
"""
from . import config, synth  # noqa: F401
from ._lib import load_library, load_plugin_library, library_path, plugin_library_path, LibraryMissing  # noqa: F401
