"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/nsvf_b200.h declares,
the ctypes table covers exactly those symbols, argument errors surface as RuntimeError, and the host-side octree
builder (no GPU involved) reproduces the reference's numbering."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from nsvf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            names += re.findall(r"NSVF_API\s+[\w\s\*]+?\b(nsvf_\w+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libnsvf_b200.so does not export " + n
    assert sorted(_lib.SIGNATURES) == names, "nsvf_b200/_lib.py SIGNATURES out of sync with include/nsvf_b200.h"
    assert _lib.load().nsvf_version() == 100


def test_every_declaration_cites_the_reference():
    src = open(os.path.join(ROOT, "include", "nsvf_b200.h")).read()
    for ref in ("intersect.cpp:49-75", "intersect.cpp:84-112", "sample.cpp:58-95", "sample.cpp:23-55",
                "octree.cpp:125-135", "encoder.py:582-590", "renderer.py:193-218", "renderer.py:88-100",
                "geometry.py:250-274", "encoder.py:605-654"):
        assert ref in src, "include/nsvf_b200.h lost its citation of " + ref


def test_no_cpu_fallback():
    """The product path refuses CPU tensors loudly (there is no CPU / PyTorch fallback)."""
    from nsvf_b200 import ops
    from nsvf_b200.clib import _ext
    x = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        _ext.aabb_intersect(x, x, torch.zeros(1, 8, 3), 0.25, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.composite(torch.zeros(2, 3), torch.zeros(2, 3, 3), torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="must be an int tensor"):
        f = torch.zeros(1, 2, 3)
        _ext.inverse_cdf_sampling(f.clone(), f, f, f, f, torch.zeros(1, 2), -1.0)     # pts_idx is float, not int
    from nsvf_b200.field import RadianceField
    with pytest.raises(RuntimeError, match="CUDA"):        # the field's fused passes have no CPU path either
        RadianceField()({"emb": torch.zeros(3, 32), "ray": torch.tensor([[0., 0., 1.]] * 3)})
    import nsvf_b200
    src = "".join(open(os.path.join(os.path.dirname(nsvf_b200.__file__), f)).read()
                  for f in os.listdir(os.path.dirname(nsvf_b200.__file__)) if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src, "the product must never import the oracle"


def test_octree_builder_matches_reference_numbering():
    from nsvf_b200.clib import _ext
    import oracle
    z = np.load(os.path.join(ROOT, "tests", "golden", "cpu_octree.npz"))
    coords = torch.from_numpy(z["coords"].astype(np.int64))
    depth = int(torch.log2((coords.max(0)[0] - coords.min(0)[0]).max().float()).ceil_().long() - 1)
    center = (coords.max(0)[0] + coords.min(0)[0]) / 2
    centers, children = _ext.build_octree(center, coords, depth)
    assert centers.dtype == torch.int32 and children.dtype == torch.int32
    assert np.array_equal(children.numpy(), z["children"])
    oc, och = oracle.build_octree(center.numpy(), coords.numpy(), depth)
    assert np.array_equal(centers.numpy(), oc) and np.array_equal(children.numpy(), och)
    # edge cases: a single depth-0 node with 8 leaves; duplicate points (a later leaf replaces the earlier one)
    pts = torch.tensor([[a, b, c] for a in (0, 2) for b in (0, 2) for c in (0, 2)])
    c1, ch1 = _ext.build_octree(torch.tensor([1., 1., 1.]), pts, 0)
    assert c1.shape == (9, 3) and ch1[-1, 8] == 2 and sorted(ch1[-1, :8].tolist()) == list(range(8))
    dup = torch.tensor([[0, 0, 0], [0, 0, 0], [2, 2, 2]])
    c2, ch2 = _ext.build_octree(torch.tensor([1., 1., 1.]), dup, 0)
    o2, oh2 = oracle.build_octree(np.array([1., 1., 1.]), dup.numpy(), 0)
    assert np.array_equal(ch2.numpy(), oh2)


def test_geometry_mirror_matches_reference_fixture():
    from nsvf_b200 import geometry
    z = np.load(os.path.join(ROOT, "tests", "golden", "cpu_kat1_encoder.npz"))
    feats, keys = geometry.corner_keys(torch.from_numpy(z["points"]), float(z["voxel_size"]) * .5)
    assert np.array_equal(feats.numpy(), z["feats"]) and np.array_equal(keys.numpy(), z["keys"])
    off = geometry.offset_points(torch.zeros(1, 3), 1.0).reshape(-1, 3)
    assert off.tolist() == [[a, b, c] for a in (-1., 1.) for b in (-1., 1.) for c in (-1., 1.)]
    lat = geometry.offset_points(torch.zeros(1, 3), 1.0, bits=16).reshape(-1, 3)
    assert lat.shape == (4096, 3) and float(lat.min()) == -1.0 and float(lat.max()) == 1.0
