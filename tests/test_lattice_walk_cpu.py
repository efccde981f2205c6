"""CPU check of the lattice walk's central claim (csrc/voxel_grid.cu): the cells a ray visits — layers of
its dominant axis, per layer the cells its span covers in the two minor axes, widened by eps — contain EVERY voxel the
reference's scan reports (the oracle, fairnr/clib/src/intersect_gpu.cu:125-167 restated in C), and visiting them in
travel order yields the hits sorted by entry depth for ordinary camera rays.  The model below restates the kernel's
enumeration in float32 numpy (same formulas, same margins; an algorithm check, not a bit-level one — the GPU tests
compare the kernel itself with the reference kernels)."""
import numpy as np
import pytest

import oracle
from nsvf_b200 import synthetic

f32 = np.float32


def _candidates(o, d, g, dims, vs):
    """Cells visited by grid_walk_kernel for ray (o, d) on the lattice with origin g (centre of cell 0), extent dims,
    cell size vs; None when the kernel would not trust the walk and scans all voxels instead."""
    u = [f32(f32(f32(o[a]) - g[a]) / f32(vs)) + f32(0.5) for a in range(3)]
    ad = [abs(f32(x)) for x in d]
    am, far = max(ad), max(abs(x) for x in u)
    if not (np.isfinite(o).all() and np.isfinite(d).all() and 1e-18 <= am <= 1e18 and far <= 1e5):
        return None
    m = (2 if ad[2] > ad[1] else 1) if ad[1] > ad[0] else (2 if ad[2] > ad[0] else 0)
    p = 0 if m == 2 else m + 1
    q = 0 if p == 2 else p + 1
    um0, up0, uq0 = u[m], u[p], u[q]
    dm, dp, dq = f32(d[m]), f32(d[p]), f32(d[q])
    nm, np_, nq = dims[m], dims[p], dims[q]
    inv = f32(1.0) / dm
    sp, sq = f32(dp * inv), f32(dq * inv)
    eps = f32(4e-3) + f32(4e-6) * f32(abs(um0) + abs(up0) + abs(uq0) + nm + np_ + nq)
    fwd = dm > 0
    lo_u, hi_u = f32(-1.0), f32(nm + 1.0)
    for s, u0, n in ((sp, up0, np_), (sq, uq0, nq)):
        if abs(s) >= 1e-6:
            t0 = um0 + f32((f32(-2.0) - u0) / s)
            t1 = um0 + f32((f32(n + 2.0) - u0) / s)
            lo_u, hi_u = max(lo_u, min(t0, t1)), min(hi_u, max(t0, t1))
        elif u0 < -4 or u0 > n + 4:
            hi_u = f32(-2.0)
    out = []
    if not (lo_u <= hi_u):
        return out
    i_lo, i_hi = max(0, int(np.floor(lo_u)) - 1), min(nm - 1, int(np.floor(hi_u)) + 1)
    i = max(i_lo, int(np.floor(um0 - eps))) if fwd else min(i_hi, int(np.floor(um0 + eps)))
    while (i <= i_hi) if fwd else (i >= i_lo):
        ua, ub = f32(i) - eps, f32(i + 1) + eps
        if fwd:
            ua = max(ua, um0 - eps)
        else:
            ub = min(ub, um0 + eps)
        ra, rb = f32(ua - um0), f32(ub - um0)
        pa, pb = f32(sp * ra + up0), f32(sp * rb + up0)
        jp0, jp1 = max(0, int(np.floor(min(pa, pb) - eps))), min(np_ - 1, int(np.floor(max(pa, pb) + eps)))
        qa, qb = f32(sq * ra + uq0), f32(sq * rb + uq0)
        jq0, jq1 = max(0, int(np.floor(min(qa, qb) - eps))), min(nq - 1, int(np.floor(max(qa, qb) + eps)))
        if jp0 <= jp1 and jq0 <= jq1:
            for a in range(jp1 - jp0 + 1):
                jp = jp0 + a if dp >= 0 else jp1 - a
                for b in range(jq1 - jq0 + 1):
                    jq = jq0 + b if dq >= 0 else jq1 - b
                    c = [0, 0, 0]
                    c[m], c[p], c[q] = i, jp, jq
                    out.append(tuple(c))
        i += 1 if fwd else -1
    return out


@pytest.mark.parametrize("name,n_rays", [("C2", 400), ("C3", 120)])
def test_walk_candidates_cover_every_hit_of_the_reference_scan(name, n_rays):
    rng = np.random.default_rng(0)
    scene = synthetic.make_scene(name)
    pts = scene.points.astype(np.float32)
    vs = f32(scene.voxel_size)
    g = pts.min(0)
    qi = np.rint((pts - g) / vs).astype(np.int64)
    assert np.abs((pts - g) - qi * vs).max() <= 1e-3 * vs           # the kernel's lattice test
    dims = (qi.max(0) + 1).tolist()
    grid = -np.ones(dims, np.int64)
    grid[qi[:, 0], qi[:, 1], qi[:, 2]] = np.arange(len(pts))
    o = rng.normal(size=(n_rays, 3)).astype(np.float32)
    o = o / np.linalg.norm(o, axis=1, keepdims=True) * 4.5
    d = rng.uniform(-1, 1, size=(n_rays, 3)).astype(np.float32) - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = n_rays // 4          # awkward rays: from inside, from cell corners, along lattice edges, zero components
    o[:k] = rng.uniform(-0.5, 0.5, size=(k, 3)).astype(np.float32)
    o[k:2 * k] = (pts[rng.integers(0, len(pts), k)] + vs * 0.5).astype(np.float32)
    d[k:k + k // 2] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k // 2)] * rng.choice([-1, 1], (k // 2, 1)).astype(np.float32)
    d[2 * k:2 * k + 10, 1] = 0
    idx, dmin, _ = oracle.aabb_intersect(o[None], d[None], pts[None], float(vs), 400)
    idx, dmin = idx[0], dmin[0]
    unsorted_plain, walked = 0, 0
    for r in range(n_rays):
        cand = _candidates(o[r], d[r], g, dims, vs)
        if cand is None:
            continue
        walked += 1
        visited = [int(grid[c]) for c in cand if grid[c] >= 0]
        assert len(set(visited)) == len(visited), "a cell was visited twice"
        hits = idx[r][idx[r] >= 0]
        assert not (set(hits.tolist()) - set(visited)), "ray %d: a hit of the reference scan is not among the visited cells" % r
        key = {int(v): (float(t), int(v)) for v, t in zip(idx[r], dmin[r]) if v >= 0}
        seq = [key[v] for v in visited if v in key]
        if r >= 2 * k + 10 and seq != sorted(seq):
            unsorted_plain += 1
    assert walked >= n_rays - 1 and int((idx >= 0).sum()) > n_rays
    assert unsorted_plain == 0, "ordinary camera rays must come out of the walk already sorted by (depth, index)"
