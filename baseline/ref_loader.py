"""Import the UNMODIFIED reference (baseline/_ref, see install_ref.py) in a process that has no fairseq.

BASELINE INFRASTRUCTURE ONLY.  fairseq / plyfile / imageio / skimage / pylab are not installed and are stubbed with
the few names the reference touches at import time; `fairnr` is registered as a namespace stub so that
fairnr/__init__.py (which pulls tasks/criterions -> the fairseq trainer) never executes, while fairnr.clib,
fairnr.data.geometry, fairnr.modules.* and fairnr.models.* are the reference's own files, executed as they are.
`fairnr.clib._ext` is the extension the reference's own setup.py built (baseline/_ref/fairnr/clib/_ext*.so).

NOTE: importing fairnr.modules.reader runs `torch.autograd.set_detect_anomaly(True)` (reader.py:12) — a process-wide
switch.  bench.py therefore runs the reference legs in a child process.
"""
import contextlib
import importlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_ROOT = os.path.join(HERE, "_ref")


def available(root=DEFAULT_ROOT):
    return os.path.exists(os.path.join(root, "fairnr", "modules", "renderer.py"))


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


class _Any:
    def __init__(self, *a, **k):
        pass


class _BaseFairseqModel(torch.nn.Module):
    """The two hooks fairnr_model.BaseModel calls on its fairseq base class (fairseq/models/fairseq_model.py)."""

    def set_num_updates(self, num_updates):
        pass

    def upgrade_state_dict_named(self, state_dict, name):
        pass


def load(root=DEFAULT_ROOT, ext=None, models=True):
    """-> namespace with clib, geometry, encoder, renderer, field, reader (and nsvf / nerf model modules)."""
    if getattr(sys.modules.get("fairnr"), "_nsvf_ref_root", None) == root:
        return _collect(models)
    deco = lambda *a, **k: (lambda x: x)

    @contextlib.contextmanager
    def with_torch_seed(seed):               # fairseq.utils.with_torch_seed: seed inside, restore outside
        st = torch.random.get_rng_state()
        cst = torch.cuda.get_rng_state() if torch.cuda.is_available() else None
        torch.manual_seed(seed)
        try:
            yield
        finally:
            torch.random.set_rng_state(st)
            if cst is not None:
                torch.cuda.set_rng_state(cst)

    _stub("fairseq").__path__ = []
    _stub("fairseq.utils", get_activation_fn=lambda n: None, with_torch_seed=with_torch_seed,
          item=lambda x: x.item() if hasattr(x, "item") else x)
    _stub("fairseq.modules", LayerNorm=torch.nn.LayerNorm)
    _stub("fairseq.meters", StopwatchMeter=_Any, TimeMeter=_Any)
    _stub("fairseq.distributed_utils", get_rank=lambda: 0, get_world_size=lambda: 1)
    _stub("fairseq.data", FairseqDataset=object, BaseWrapperDataset=object)
    _stub("fairseq.models", BaseFairseqModel=_BaseFairseqModel, register_model=deco, register_model_architecture=deco)
    _stub("plyfile", PlyData=_Any, PlyElement=_Any)
    _stub("imageio")
    _stub("skimage").__path__ = []
    _stub("skimage.metrics")
    _stub("pylab")
    pkg = _stub("fairnr")
    pkg.__path__ = [os.path.join(root, "fairnr")]
    pkg._nsvf_ref_root = root
    pkg.ResetTrainerException = type("ResetTrainerException", (Exception,), {})
    if ext is not None:
        sys.modules["fairnr.clib._ext"] = ext
    clib = importlib.import_module("fairnr.clib")
    if ext is not None:
        clib._ext = ext
    return _collect(models)


def _collect(models):
    ns = types.SimpleNamespace()
    ns.clib = importlib.import_module("fairnr.clib")
    ns.geometry = importlib.import_module("fairnr.data.geometry")
    ns.encoder = importlib.import_module("fairnr.modules.encoder")
    ns.renderer = importlib.import_module("fairnr.modules.renderer")
    ns.field = importlib.import_module("fairnr.modules.field")
    ns.reader = importlib.import_module("fairnr.modules.reader")
    if models:
        ns.nerf = importlib.import_module("fairnr.models.nerf")
        ns.nsvf = importlib.import_module("fairnr.models.nsvf")
    return ns
