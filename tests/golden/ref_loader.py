"""Import the UNMODIFIED reference Python modules from /root/reference on CPU (test infrastructure).

fairseq / plyfile / imageio / skimage are not installed and are stubbed; `fairnr` is registered as a namespace
stub so that fairnr/__init__.py (which pulls tasks/criterions -> fairseq) never executes.  Only works where
/root/reference exists (the build container); the GPU box uses the committed fixtures instead.
"""
import contextlib
import importlib
import os
import sys
import types

import torch

REF = os.environ.get("NSVF_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "fairnr"))


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


class _Any:
    def __init__(self, *a, **k):
        pass


def load(ext=None):
    """Returns (clib, geometry, encoder, renderer, field) reference modules; `ext` is bound as fairnr.clib._ext."""
    if "fairnr.modules.encoder" in sys.modules and getattr(sys.modules["fairnr"], "_nsvf_ref", False):
        m = sys.modules
        return m["fairnr.clib"], m["fairnr.data.geometry"], m["fairnr.modules.encoder"], m["fairnr.modules.renderer"], \
            m["fairnr.modules.field"]
    deco = lambda *a, **k: (lambda x: x)

    @contextlib.contextmanager
    def with_torch_seed(seed):
        st = torch.random.get_rng_state()
        torch.manual_seed(seed)
        try:
            yield
        finally:
            torch.random.set_rng_state(st)

    _stub("fairseq").__path__ = []
    _stub("fairseq.utils", get_activation_fn=lambda n: None, with_torch_seed=with_torch_seed,
          item=lambda x: x.item() if hasattr(x, "item") else x)
    _stub("fairseq.modules", LayerNorm=torch.nn.LayerNorm)
    _stub("fairseq.meters", StopwatchMeter=_Any, TimeMeter=_Any)
    _stub("fairseq.distributed_utils", get_rank=lambda: 0, get_world_size=lambda: 1)
    _stub("fairseq.data", FairseqDataset=object, BaseWrapperDataset=object)
    _stub("fairseq.models", BaseFairseqModel=torch.nn.Module, register_model=deco, register_model_architecture=deco)
    _stub("plyfile", PlyData=_Any, PlyElement=_Any)
    _stub("imageio")
    _stub("skimage").__path__ = []
    _stub("skimage.metrics")
    _stub("pylab")
    pkg = _stub("fairnr")
    pkg.__path__ = [os.path.join(REF, "fairnr")]
    pkg._nsvf_ref = True
    pkg.ResetTrainerException = type("ResetTrainerException", (Exception,), {})
    if ext is None:
        from oracle import build_ref
        ext = build_ref.load()
    if ext is not None:
        sys.modules["fairnr.clib._ext"] = ext
    clib = importlib.import_module("fairnr.clib")
    if ext is not None:
        clib._ext = ext
    geo = importlib.import_module("fairnr.data.geometry")
    enc = importlib.import_module("fairnr.modules.encoder")
    ren = importlib.import_module("fairnr.modules.renderer")
    fld = importlib.import_module("fairnr.modules.field")
    return clib, geo, enc, ren, fld
