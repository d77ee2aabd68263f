"""Import the REAL reference modules from /root/reference (this container only).

TEST INFRASTRUCTURE. Used by `oracle/make_golden.py` (to produce tests/golden/*.npz) and by the
`not gpu` tests that pin the oracle against the reference when /root/reference is mounted.
Nothing in the product path, in `-m gpu` tests, in smoke() or bench.py may import this file:
/root/reference does not exist on the GPU box.

Recipe: the reference packages cannot be imported as shipped
(CLAP/src/laion_clap/__init__.py pulls librosa/wget/h5py/webdataset and does HF downloads at
module import, training/data.py:44-46), so
  * `torchlibrosa` / `h5py` are satisfied by the stand-ins under oracle/_stubs,
  * `clap_module` is registered as a bare namespace package so its __init__ is skipped,
  * `get_audio_features` & the int16 helpers are AST-lifted out of training/data.py,
  * `src.residual` is imported with a stub `CLAP` module that exposes the lifted function.
"""
import ast
import importlib
import json
import os
import sys
import types
from contextlib import suppress

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference when mounted (build container); else the snapshot oracle/build_ref.py made of the same files (oracle/_ref,
# git-ignored, shipped to the GPU box) so bench.py --impl reference can time the reference's own modules there
REF = os.environ.get("ARD_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/CLAP") else os.path.join(_HERE, "_ref"))
_STUBS = os.path.join(_HERE, "_stubs")


def available():
    return os.path.isdir(os.path.join(REF, "CLAP", "src", "laion_clap", "clap_module"))


_cache = {}


def load():
    """Returns a namespace with: htsat, model (modules), residual (src.residual), get_audio_features,
    float32_to_int16, int16_to_float32, cfg(name) -> model config dict."""
    if _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REF)
    import numpy as np
    import torch
    import torch.nn.functional as F
    import torchaudio
    import torchvision

    if _STUBS not in sys.path:
        sys.path.insert(0, _STUBS)
    cm_dir = os.path.join(REF, "CLAP", "src", "laion_clap", "clap_module")
    pkg = types.ModuleType("clap_module")
    pkg.__path__ = [cm_dir]
    sys.modules["clap_module"] = pkg
    htsat = importlib.import_module("clap_module.htsat")
    model = importlib.import_module("clap_module.model")

    # AST-lift the featuriser (training/data.py:93-108, 363-506)
    data_py = os.path.join(REF, "CLAP", "src", "laion_clap", "training", "data.py")
    tree = ast.parse(open(data_py).read())
    want = {"get_mel", "get_audio_features", "int16_to_float32", "float32_to_int16",
            "int16_to_float32_torch", "float32_to_int16_torch"}
    g = {"torch": torch, "np": np, "torchaudio": torchaudio, "torchvision": torchvision, "F": F,
         "suppress": suppress}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            exec(compile(ast.Module([node], []), data_py, "exec"), g)

    stub = types.ModuleType("CLAP")
    stub.get_audio_features = g["get_audio_features"]
    stub.int16_to_float32 = g["int16_to_float32"]
    stub.float32_to_int16 = g["float32_to_int16"]
    sys.modules["CLAP"] = stub
    src = types.ModuleType("src")
    src.__path__ = [os.path.join(REF, "src")]
    sys.modules["src"] = src
    residual = importlib.import_module("src.residual")

    def cfg(name):
        c = json.load(open(os.path.join(cm_dir, "model_configs", "HTSAT-%s.json" % name)))
        c["text_cfg"]["model_type"] = "transformer"  # avoids RobertaModel.from_pretrained (model.py:505-506)
        return c

    ns = types.SimpleNamespace(htsat=htsat, model=model, residual=residual,
                               get_audio_features=g["get_audio_features"], get_mel=g["get_mel"],
                               float32_to_int16=g["float32_to_int16"], int16_to_float32=g["int16_to_float32"],
                               cfg=cfg)
    _cache["ns"] = ns
    return ns


def build_clap(name="tiny", enable_fusion=False, fusion_type="None"):
    """clap_module.model.CLAP with the audio branch of HTSAT-<name> (text tower = small unused transformer)."""
    ns = load()
    c = ns.cfg(name)
    clap = ns.model.CLAP(**c, enable_fusion=enable_fusion, fusion_type=fusion_type).eval()
    return clap, c
