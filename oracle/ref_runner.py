"""Run the reference's UNMODIFIED Inference_*.py scripts under a harness (TEST INFRASTRUCTURE).

The scripts are top-level code that builds an `onnxruntime.InferenceSession`, loads audio with
pydub, loops over chunks and writes two timestamp files.  onnxruntime / pydub are not installed
here, so this module puts fakes into sys.modules:
  * onnxruntime.InferenceSession(path, ...) -> `session_factory(path)`: an object with the ORT
    surface the scripts touch (run / get_inputs / get_outputs / get_providers / _inputs_meta /
    _outputs_meta) that executes the reference's own PyTorch wrapper module (AST-loaded from the
    Export script) with our seeded weights;
  * pydub.AudioSegment.from_file(p).set_channels(1).set_frame_rate(sr).get_array_of_samples()
    -> the same wave + audioop.tomono + audioop.ratecv calls pydub makes.
Module-level UPPER_CASE constants can be overridden (the reference's configuration mechanism).
The script runs in a scratch directory, so its timestamp files land there and are returned.
Only usable where /root/reference exists; never imported by the product.
"""
from __future__ import annotations

import ast
import contextlib
import importlib.machinery
import io
import os
import sys
import tempfile
import types

import numpy as np

from . import ref_loader as RL


class NodeArg:
    def __init__(self, name, shape, type_):
        self.name, self.shape, self.type = name, shape, type_


class FakeSession:
    """ORT-shaped wrapper around a torch module.  `fn(feed: dict[str, np.ndarray]) -> list[np.ndarray]`"""

    def __init__(self, inputs, outputs, fn):
        self._inputs_meta, self._outputs_meta, self._fn = inputs, outputs, fn
        self.calls = []

    def get_inputs(self):
        return self._inputs_meta

    def get_outputs(self):
        return self._outputs_meta

    def get_providers(self):
        return ["CPUExecutionProvider"]

    def run(self, output_names, feed, run_options=None):
        outs = self._fn(feed)
        self.calls.append(({k: np.array(v, copy=True) for k, v in feed.items()}, [np.array(o, copy=True) for o in outs]))
        names = [o.name for o in self._outputs_meta]
        if output_names is None:
            return outs
        return [outs[names.index(n)] for n in output_names]


def _fake_onnxruntime(session_factory):
    m = types.ModuleType("onnxruntime")
    m.__spec__ = importlib.machinery.ModuleSpec("onnxruntime", loader=None)

    class SessionOptions:
        def add_session_config_entry(self, *_a):
            pass

        def add_run_config_entry(self, *_a):
            pass

    m.SessionOptions = SessionOptions
    m.ExecutionMode = types.SimpleNamespace(ORT_SEQUENTIAL=0, ORT_PARALLEL=1)
    m.GraphOptimizationLevel = types.SimpleNamespace(ORT_ENABLE_ALL=99, ORT_DISABLE_ALL=0, ORT_ENABLE_BASIC=1,
                                                     ORT_ENABLE_EXTENDED=2)
    m.RunOptions = SessionOptions
    m.get_available_providers = lambda: ["CPUExecutionProvider"]
    m.InferenceSession = lambda path, *a, **k: session_factory(path)
    m.OrtValue = types.SimpleNamespace()
    return m


def _fake_pydub(load_wav):
    m = types.ModuleType("pydub")
    m.__spec__ = importlib.machinery.ModuleSpec("pydub", loader=None)

    class _Seg:
        def __init__(self, path):
            self.path, self.sr = path, None

        def set_channels(self, n):
            assert n == 1
            return self

        def set_frame_rate(self, sr):
            self.sr = sr
            return self

        def get_array_of_samples(self):
            return load_wav(self.path, self.sr)

    class AudioSegment:
        @staticmethod
        def from_file(path, *a, **k):
            return _Seg(path)

    m.AudioSegment = AudioSegment
    return m


def run_script(rel_path: str, session_factory, load_wav, overrides: dict | None = None, seed: int = 0,
               files_to_link: dict | None = None, extra_modules: dict | None = None):
    """exec the reference script; returns (namespace, {filename: text written in the scratch dir})."""
    full = os.path.join(RL.REF_ROOT, rel_path)
    tree = ast.parse(open(full, encoding="utf-8").read(), filename=full)
    overrides = dict(overrides or {})
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if name in overrides:
                node.value = ast.Constant(overrides.pop(name))
    if overrides:
        raise KeyError(f"constants not found in {rel_path}: {sorted(overrides)}")
    ast.fix_missing_locations(tree)
    code = compile(tree, full, "exec")
    saved = {k: sys.modules.get(k) for k in ("onnxruntime", "pydub")}
    sys.modules["onnxruntime"] = _fake_onnxruntime(session_factory)
    sys.modules["pydub"] = _fake_pydub(load_wav)
    for k, v in (extra_modules or {}).items():
        saved[k] = sys.modules.get(k)
        sys.modules[k] = v
    cwd = os.getcwd()
    ns = {"__name__": "__main__", "__file__": full}
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, src in (files_to_link or {}).items():
            os.symlink(src, os.path.join(tmp, name))
        os.chdir(tmp)
        try:
            np.random.seed(seed)
            with contextlib.redirect_stdout(io.StringIO()) as log:
                exec(code, ns)
            ns["__stdout__"] = log.getvalue()
            for fn in os.listdir(tmp):
                p = os.path.join(tmp, fn)
                if os.path.isfile(p) and not os.path.islink(p):
                    out[fn] = open(p, encoding="utf-8").read()
        finally:
            os.chdir(cwd)
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    return ns, out
