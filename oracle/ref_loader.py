"""Load classes/functions out of the read-only reference checkout WITHOUT copying them.

Test infrastructure; only usable where /root/reference exists (the authoring
container).  The reference's Export_*/Inference_* files are scripts that run
on import, so we parse them with `ast`, keep only ClassDef / FunctionDef /
simple constant assignments and exec those nodes in a fresh namespace.
Missing third-party modules (onnxruntime, onnxslim, pydub, funasr, ...) are
replaced by inert stubs that carry a __spec__ (torch._dynamo inspects it).
"""
from __future__ import annotations

import ast
import importlib.machinery
import importlib.util
import math
import os
import sys
import types

REF_ROOT = os.environ.get("VADX_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(REF_ROOT)


def _stub(name: str) -> types.ModuleType:
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []  # behave like a package so "a.b" imports resolve
    sys.modules[name] = m
    return m


def install_stubs() -> None:
    for n in ("onnxruntime", "onnxslim", "pydub", "kaldiio"):
        try:
            importlib.import_module(n)
        except Exception:
            _stub(n)
    ort = sys.modules["onnxruntime"]
    if not hasattr(ort, "InferenceSession"):
        ort.InferenceSession = object
    slim = sys.modules["onnxslim"]
    if not hasattr(slim, "slim"):
        slim.slim = lambda *a, **k: None
    pd = sys.modules["pydub"]
    if not hasattr(pd, "AudioSegment"):
        pd.AudioSegment = object
    # funasr.register.tables.register(...) is used as a class decorator
    try:
        importlib.import_module("funasr.register")
    except Exception:
        _stub("funasr")
        reg = _stub("funasr.register")

        class _Tables:
            @staticmethod
            def register(*_a, **_k):
                return lambda cls: cls

        reg.tables = _Tables()
        sys.modules["funasr"].register = reg


def import_file(path: str, modname: str):
    """Plain import of a reference file that is import-safe (STFT_Process.py, encoder.py)."""
    install_stubs()
    full = os.path.join(REF_ROOT, path)
    spec = importlib.util.spec_from_file_location(modname, full)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def _is_simple_const(node: ast.AST) -> bool:
    """UPPER_CASE / _UPPER = <expression without calls to heavy stuff>."""
    if not isinstance(node, ast.Assign) or len(node.targets) != 1:
        return False
    t = node.targets[0]
    if not isinstance(t, ast.Name):
        return False
    name = t.id
    return name.upper() == name and any(c.isalpha() for c in name)


def extract(path: str, extra_ns: dict | None = None, want_consts: bool = True) -> dict:
    """exec only the definitions of a reference *script*; return the namespace."""
    install_stubs()
    import numpy as np
    import torch

    full = os.path.join(REF_ROOT, path)
    with open(full, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=full)
    keep = []
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)):
            keep.append(node)
        elif want_consts and _is_simple_const(node):
            keep.append(node)
    ns: dict = {"torch": torch, "np": np, "math": math, "F": torch.nn.functional,
                "__name__": "ref_extract_" + os.path.basename(path)}
    try:
        import torchaudio
        ns["torchaudio"] = torchaudio
    except Exception:
        pass
    from datetime import timedelta
    ns["timedelta"] = timedelta
    if extra_ns:
        ns.update(extra_ns)
    for node in keep:
        mod = ast.Module(body=[node], type_ignores=[])
        try:
            exec(compile(mod, full, "exec"), ns)
        except Exception:
            # constants that depend on things we did not keep (paths, sessions) are skipped
            if isinstance(node, (ast.ClassDef, ast.FunctionDef)):
                raise
    return ns


class Args:
    """Stand-in for the `package["args"]` namespace of a FireRed checkpoint."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
