"""CPU oracle for the NSVF hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this package.  The product (nsvf_b200/) never imports it and has no CPU path.

  * nsvf_oracle.c      C restatement of the reference kernels and torch-level stages (numpy in / out here)
  * wrappers.py        numpy restatement of the reference's Python glue (clib/__init__.py tiling,
                       encoder.ray_intersect sort, nsvf.py probs/steps, splitting_points, pruning)
  * build_ref.py       compiles the UNMODIFIED reference extension into oracle/_ref/ref_ext.so
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRC = os.path.join(_HERE, "nsvf_oracle.c")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC",
                               "-fvisibility=hidden", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_octree_build.restype = ctypes.c_longlong
        _lib.oracle_svo_intersect.restype = ctypes.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


c_int, c_ll, c_float = ctypes.c_int, ctypes.c_longlong, ctypes.c_float


def aabb_intersect(ray_start, ray_dir, points, voxelsize, n_max, inv_dir=None):
    """reference aabb_intersect (intersect.cpp:49-75): rays [B,M,3], points [B,n,3] or [n,3] (shared)."""
    rs, rd = _f(ray_start), _f(ray_dir)
    pts = _f(points)
    b, m = rs.shape[:2]
    shared = pts.ndim == 2
    n = pts.shape[-2]
    inv = None if inv_dir is None else _f(inv_dir)
    idx = np.empty((b, m, n_max), np.int32)
    dmin = np.empty((b, m, n_max), np.float32)
    dmax = np.empty((b, m, n_max), np.float32)
    lib().oracle_aabb_intersect(c_int(b), c_int(n), c_int(m), c_float(voxelsize), c_int(n_max), _ptr(rs), _ptr(rd),
                                _ptr(inv), _ptr(pts), c_ll(0 if shared else n * 3), _ptr(idx), _ptr(dmin), _ptr(dmax))
    return idx, dmin, dmax


def svo_intersect(ray_start, ray_dir, points, children, voxelsize, n_max, inv_dir=None):
    rs, rd = _f(ray_start), _f(ray_dir)
    pts, ch = _f(points), _i(children)
    b, m = rs.shape[:2]
    shared = pts.ndim == 2
    T = pts.shape[-2]
    inv = None if inv_dir is None else _f(inv_dir)
    idx = np.empty((b, m, n_max), np.int32)
    dmin = np.empty((b, m, n_max), np.float32)
    dmax = np.empty((b, m, n_max), np.float32)
    lib().oracle_svo_intersect(c_int(b), c_int(T), c_int(m), c_float(voxelsize), c_int(n_max), _ptr(rs), _ptr(rd),
                               _ptr(inv), _ptr(pts), _ptr(ch), c_ll(0 if shared else T), _ptr(idx), _ptr(dmin),
                               _ptr(dmax))
    return idx, dmin, dmax


def uniform_ray_sampling(pts_idx, min_depth, max_depth, noise, step_size, max_steps):
    pi, mn, mx, nz = _i(pts_idx), _f(min_depth), _f(max_depth), _f(noise)
    g, r, p = mn.shape
    si = np.empty((g, r, max_steps), np.int32)
    sd = np.empty((g, r, max_steps), np.float32)
    ss = np.empty((g, r, max_steps), np.float32)
    lib().oracle_uniform_ray_sampling(c_int(g), c_int(r), c_int(p), c_int(max_steps), c_float(step_size), _ptr(pi),
                                      _ptr(mn), _ptr(mx), _ptr(nz), _ptr(si), _ptr(sd), _ptr(ss))
    return si, sd, ss


def inverse_cdf_sampling(pts_idx, min_depth, max_depth, noise, probs, steps, fixed_step_size):
    pi, mn, mx, nz, pr, st = _i(pts_idx), _f(min_depth), _f(max_depth), _f(noise), _f(probs), _f(steps)
    g, r, p = mn.shape
    max_steps = nz.shape[-1]
    si = np.empty((g, r, max_steps), np.int32)
    sd = np.empty((g, r, max_steps), np.float32)
    ss = np.empty((g, r, max_steps), np.float32)
    lib().oracle_inverse_cdf_sampling(c_int(g), c_int(r), c_int(p), c_int(max_steps), c_float(fixed_step_size),
                                      _ptr(pi), _ptr(mn), _ptr(mx), _ptr(nz), _ptr(pr), _ptr(st), _ptr(si), _ptr(sd),
                                      _ptr(ss))
    return si, sd, ss


def build_octree(center, points, depth):
    c = _f(center)
    p = np.ascontiguousarray(points, dtype=np.int64)
    T = lib().oracle_octree_build(_ptr(c), _ptr(p), c_ll(p.shape[0]), c_int(int(depth)))
    centers = np.empty((T, 3), np.int32)
    children = np.empty((T, 9), np.int32)
    lib().oracle_octree_flatten(_ptr(centers), _ptr(children))
    return centers, children


def trilinear_fwd(sampled_idx, xyz, feats, centres, values, voxel_size):
    si, x, ft, c, v = _i(sampled_idx), _f(xyz), _i(feats), _f(centres), _f(values)
    M, D = si.shape[0], v.shape[1]
    out = np.empty((M, D), np.float32)
    lib().oracle_trilinear_fwd(c_ll(M), c_int(D), _ptr(si), _ptr(x), _ptr(ft), _ptr(c), _ptr(v), c_float(voxel_size),
                               _ptr(out))
    return out


def trilinear_bwd(sampled_idx, xyz, feats, centres, values, voxel_size, grad_out):
    si, x, ft, c, v, g = _i(sampled_idx), _f(xyz), _i(feats), _f(centres), _f(values), _f(grad_out)
    M, (Kc, D) = si.shape[0], v.shape
    gv = np.empty((Kc, D), np.float32)
    gx = np.empty((M, 3), np.float32)
    lib().oracle_trilinear_bwd(c_ll(M), c_int(D), c_ll(Kc), _ptr(si), _ptr(x), _ptr(ft), _ptr(c), _ptr(v),
                               c_float(voxel_size), _ptr(g), _ptr(gv), _ptr(gx))
    return gv, gx


def composite_fwd(fe, tex, depth):
    fe, dp = _f(fe), _f(depth)
    tex = None if tex is None else _f(tex)
    B, K = fe.shape
    probs = np.empty((B, K), np.float32)
    od = np.empty((B,), np.float32)
    om = np.empty((B,), np.float32)
    oc = np.zeros((B, 3), np.float32)
    lib().oracle_composite_fwd(c_ll(B), c_int(K), _ptr(fe), _ptr(tex), _ptr(dp), _ptr(probs), _ptr(od), _ptr(om),
                               _ptr(oc))
    return probs, od, om, oc


def composite_bwd(fe, tex, depth, g_probs, g_depth, g_missed, g_colors):
    fe, dp = _f(fe), _f(depth)
    tex = None if tex is None else _f(tex)
    gs = [None if g is None else _f(g) for g in (g_probs, g_depth, g_missed, g_colors)]
    B, K = fe.shape
    g_fe = np.empty((B, K), np.float32)
    g_tex = np.zeros((B, K, 3), np.float32)
    lib().oracle_composite_bwd(c_ll(B), c_int(K), _ptr(fe), _ptr(tex), _ptr(dp), _ptr(gs[0]), _ptr(gs[1]), _ptr(gs[2]),
                               _ptr(gs[3]), _ptr(g_fe), _ptr(g_tex if tex is not None else None))
    return g_fe, g_tex
