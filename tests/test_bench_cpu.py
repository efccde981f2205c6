"""CPU tests of the measurement contract: the reference arm of bench.py (oracle port on host cores) prints one JSON
line with the required keys, and hypothesis-driven properties of the oracle that the GPU tests rely on."""
import json
import os
import subprocess
import sys

import numpy as np
from hypothesis import given, settings, strategies as st

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-fraction", "128"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["vs_baseline"] is None
    # non-zero ranks of a torchrun launch exit 0 without work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                          capture_output=True, text=True, timeout=120, env=env)
    assert out2.returncode == 0 and out2.stdout.strip() == ""


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 40), st.integers(1, 70))
def test_composite_probabilities_sum_to_one(seed, B, K):
    rng = np.random.RandomState(seed)
    fe = (np.abs(rng.randn(B, K)) * rng.rand(B, 1)).astype(np.float32)
    depth = np.sort(rng.rand(B, K).astype(np.float32), 1)
    probs, d, missed, colors = oracle.composite_fwd(fe, np.ones((B, K, 3), np.float32), depth)
    assert np.all(probs >= 0) and np.allclose(probs.sum(1) + missed, 1, atol=2e-6)
    assert np.allclose(colors, (1 - missed)[:, None], atol=2e-6)          # constant texture 1 -> colour = opacity
    assert np.all(d <= depth.max(1) * (1 - missed) + 1e-5)


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 64), st.integers(1, 9))
def test_aabb_hits_are_first_n_max_in_index_order(seed, n_rays, n_max):
    rng = np.random.RandomState(seed)
    g = rng.randint(2, 6)
    pts = np.stack(np.meshgrid(*[np.arange(g, dtype=np.float32) * 0.5] * 3, indexing="ij"), -1).reshape(-1, 3)
    pts = pts[rng.rand(len(pts)) < 0.8]
    if len(pts) == 0:
        return
    o = rng.randn(n_rays, 3).astype(np.float32) * 3
    d = (pts[rng.randint(0, len(pts), n_rays)] + rng.randn(n_rays, 3).astype(np.float32) * 0.2 - o)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    full = oracle.aabb_intersect(o[None], d[None].astype(np.float32), pts, 0.5, len(pts))[0][0]
    cut, dmin, dmax = [a[0] for a in oracle.aabb_intersect(o[None], d[None].astype(np.float32), pts, 0.5, n_max)]
    m = min(n_max, full.shape[1])
    assert np.array_equal(cut[:, :m], full[:, :m]) and np.all(cut[:, m:] == -1)
    valid = cut >= 0
    assert np.all(dmax[valid] >= dmin[valid]) and np.all(dmin[valid] >= 0)
    for r in range(n_rays):                       # ascending voxel index, then -1 padding
        v = cut[r][cut[r] >= 0]
        assert np.all(np.diff(v) > 0) and np.all(cut[r][len(v):] == -1)
