"""Import shim: the package directory is named `s2vt-video-caption_b200` (not a Python identifier), so this
module loads it under the importable name `s2vt_b200`.

    import s2vt_b200
    model = s2vt_b200.S2VT(vocab_size, feat_dim, length, dim_hid=512, dim_embed=512).cuda()
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "s2vt-video-caption_b200")
_NAME = "s2vt_b200"

_spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[_NAME] = _mod          # replaces this shim; submodules import as s2vt_b200.<name>
_spec.loader.exec_module(_mod)
