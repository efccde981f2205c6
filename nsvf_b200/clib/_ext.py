"""Level-1 boundary: the 7 functions of the reference's pybind module `fairnr.clib._ext`
(fairnr/clib/src/binding.cpp:11-20) with the same names, argument order, dtypes, output shapes and
fill values, implemented on the sm_100a kernels behind the C ABI (include/nsvf_b200.h).

Checks mirror fairnr/clib/include/utils.h:10-30 (CHECK_CONTIGUOUS / CHECK_IS_FLOAT / CHECK_IS_INT /
CHECK_CUDA -> RuntimeError with the same message text).  Scalar arguments may be Python numbers or
0-dim tensors (the reference passes `self.voxel_size`, `self.max_hits` buffers; pybind converts them
through __float__/__int__, so max_hits = 202.5 arrives as 202).
"""
import torch

from .. import _lib

_L = _lib.load()
_p = _lib.ptr


def _chk(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _check(floats, ints=None):
    """Same order as the reference: CHECK_CONTIGUOUS on everything, then dtypes, then CHECK_CUDA."""
    ints = ints or {}
    for name, t in list(ints.items()) + list(floats.items()):
        _chk(t.is_contiguous(), name + " must be a contiguous tensor")
    for name, t in floats.items():
        _chk(t.dtype == torch.float32, name + " must be a float tensor")
    for name, t in ints.items():
        _chk(t.dtype == torch.int32, name + " must be an int tensor")
    for name, t in list(ints.items()) + list(floats.items()):
        _chk(t.is_cuda, name + " must be a CUDA tensor")


def _check_float_cuda(**tensors):
    _check(tensors)


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def _hit_outputs(ray_start, n_max):
    b, m = ray_start.shape[0], ray_start.shape[1]
    idx = torch.empty((b, m, n_max), dtype=torch.int32, device=ray_start.device)
    dmin = torch.empty((b, m, n_max), dtype=torch.float32, device=ray_start.device)
    dmax = torch.empty((b, m, n_max), dtype=torch.float32, device=ray_start.device)
    return idx, dmin, dmax


def aabb_intersect(ray_start, ray_dir, points, voxelsize, n_max, shared_points=False):
    """fairnr/clib/src/intersect.cpp:49-75.  ray_start, ray_dir f32 [B,M,3]; points f32 [B,n,3]
    -> idx i32 [B,M,n_max] (-1 fill), min_depth, max_depth f32 [B,M,n_max] (0 fill).

    `shared_points=True` (extension): points is [1,n,3] or [n,3] and serves every batch row, which is
    what the Level-2 wrapper uses instead of materialising B copies (fairnr/clib/__init__.py:71).
    """
    _check_float_cuda(ray_start=ray_start, ray_dir=ray_dir, points=points)
    voxelsize, n_max = float(voxelsize), int(n_max)
    b, m = ray_start.shape[0], ray_start.shape[1]
    if shared_points:
        n, stride, trees = points.shape[-2], 0, 1
    else:
        _chk(points.dim() == 3 and points.shape[0] == b, "points must be [B, n, 3] with B == ray_start.size(0)")
        n, stride, trees = points.shape[1], points.shape[1] * 3, b
    idx, dmin, dmax = _hit_outputs(ray_start, n_max)
    with torch.cuda.device(ray_start.device):
        ws_bytes = _L.nsvf_aabb_workspace_bytes(n, trees)
        ws = _workspace(ws_bytes, ray_start.device)
        _lib.check(_L.nsvf_aabb_intersect(
            _lib.current_stream(ray_start.device), b, n, m, voxelsize, n_max, _p(ray_start), _p(ray_dir),
            _p(points), stride, _p(idx), _p(dmin), _p(dmax), _p(ws), ws.numel()))
    return idx, dmin, dmax


def _aabb_args(ray_start, ray_dir, points, shared_points):
    _check_float_cuda(ray_start=ray_start, ray_dir=ray_dir, points=points)
    b, m = ray_start.shape[0], ray_start.shape[1]
    if shared_points:
        return b, m, points.shape[-2], 0, 1
    _chk(points.dim() == 3 and points.shape[0] == b, "points must be [B, n, 3] with B == ray_start.size(0)")
    return b, m, points.shape[1], points.shape[1] * 3, b


class AabbIndex:
    """A voxel set prepared for intersection (nsvf_aabb_prepare): the lattice / hierarchy workspace plus the centres it
    was built from.  Build it once per voxel-set change and pass it as `index=` to the queries below."""

    def __init__(self, points, voxelsize, shared_points=False):
        _check_float_cuda(points=points)
        self.points, self.voxelsize = points, float(voxelsize)
        if shared_points:
            self.n, self.stride, self.sets = points.shape[-2], 0, 1
        else:
            _chk(points.dim() == 3, "points must be [B, n, 3]")
            self.n, self.stride, self.sets = points.shape[1], points.shape[1] * 3, points.shape[0]
        with torch.cuda.device(points.device):
            self.ws = _workspace(_L.nsvf_aabb_workspace_bytes(self.n, self.sets), points.device)
            _lib.check(_L.nsvf_aabb_prepare(_lib.current_stream(points.device), self.sets, self.n, self.voxelsize,
                                            _p(points), self.stride, _p(self.ws), self.ws.numel()))

    def query(self, mode, ray_start, ray_dir, n_max, empty_depth):
        _check_float_cuda(ray_start=ray_start, ray_dir=ray_dir)
        b, m = ray_start.shape[0], ray_start.shape[1]
        _chk(self.stride == 0 or b == self.sets, "one voxel set per batch row was prepared")
        dev = ray_start.device
        idx = dmin = dmax = None
        if mode != 2:
            idx, dmin, dmax = _hit_outputs(ray_start, int(n_max))
        hits = torch.empty((b, m), dtype=torch.uint8, device=dev) if mode != 0 else None
        with torch.cuda.device(dev):
            _lib.check(_L.nsvf_aabb_intersect_prepared(
                _lib.current_stream(dev), mode, b, self.n, m, self.voxelsize, int(n_max), float(empty_depth),
                _p(ray_start), _p(ray_dir), _p(self.points), self.stride, _p(idx) if idx is not None else None,
                _p(dmin) if dmin is not None else None, _p(dmax) if dmax is not None else None,
                _p(hits) if hits is not None else None, _p(self.ws), self.ws.numel()))
        return idx, dmin, dmax, hits


def aabb_intersect_sorted(ray_start, ray_dir, points, voxelsize, n_max, empty_depth=10000.0, shared_points=False,
                          index=None):
    """Extension (not in the reference _ext): aabb_intersect + the sort / fill / any() of
    SparseVoxelEncoder.ray_intersect (encoder.py:519-524) in one kernel.
    -> idx i32, min_depth f32, max_depth f32 [B,M,n_max] sorted by entry depth, hits bool [B,M].
    `index` (AabbIndex of the same points) skips rebuilding the lattice / hierarchy."""
    if index is not None:
        idx, dmin, dmax, hits = index.query(1, ray_start, ray_dir, n_max, empty_depth)
        return idx, dmin, dmax, hits.bool()
    b, m, n, stride, trees = _aabb_args(ray_start, ray_dir, points, shared_points)
    voxelsize, n_max = float(voxelsize), int(n_max)
    idx, dmin, dmax = _hit_outputs(ray_start, n_max)
    hits = torch.empty((b, m), dtype=torch.uint8, device=ray_start.device)
    with torch.cuda.device(ray_start.device):
        ws = _workspace(_L.nsvf_aabb_workspace_bytes(n, trees), ray_start.device)
        _lib.check(_L.nsvf_aabb_intersect_sorted(
            _lib.current_stream(ray_start.device), b, n, m, voxelsize, n_max, float(empty_depth), _p(ray_start),
            _p(ray_dir), _p(points), stride, _p(idx), _p(dmin), _p(dmax), _p(hits), _p(ws), ws.numel()))
    return idx, dmin, dmax, hits.bool()


def sort_hits_by_depth(idx, min_depth, max_depth, empty_depth=10000.0):
    """Extension: in-place masked_fill + sort-by-entry-depth + any() of encoder.py:519-524 on [.., n_max] hit lists.
    Returns hits bool [..]."""
    _check(dict(min_depth=min_depth, max_depth=max_depth), dict(idx=idx))
    n_max = idx.shape[-1]
    rays = idx.numel() // max(n_max, 1)
    hits = torch.empty(idx.shape[:-1], dtype=torch.uint8, device=idx.device)
    with torch.cuda.device(idx.device):
        _lib.check(_L.nsvf_sort_hits_by_depth(_lib.current_stream(idx.device), rays, n_max, float(empty_depth),
                                              _p(idx), _p(min_depth), _p(max_depth), _p(hits)))
    return hits.bool()


def aabb_hit_mask(ray_start, ray_dir, points, voxelsize, shared_points=False, index=None):
    """Extension: hits bool [B,M] = any(aabb_intersect(...).idx != -1) without producing the hit lists."""
    if index is not None:
        return index.query(2, ray_start, ray_dir, 0, 0.0)[3].bool()
    b, m, n, stride, trees = _aabb_args(ray_start, ray_dir, points, shared_points)
    hits = torch.empty((b, m), dtype=torch.uint8, device=ray_start.device)
    with torch.cuda.device(ray_start.device):
        ws = _workspace(_L.nsvf_aabb_workspace_bytes(n, trees), ray_start.device)
        _lib.check(_L.nsvf_aabb_hit_mask(
            _lib.current_stream(ray_start.device), b, n, m, float(voxelsize), _p(ray_start), _p(ray_dir), _p(points),
            stride, _p(hits), _p(ws), ws.numel()))
    return hits.bool()


def svo_intersect(ray_start, ray_dir, points, children, voxelsize, n_max, shared_tree=False):
    """fairnr/clib/src/intersect.cpp:84-112.  points f32 [B,T,3] node centres, children i32 [B,T,9]."""
    _check_float_cuda(ray_start=ray_start, ray_dir=ray_dir, points=points)
    _chk(children.is_contiguous(), "children must be a contiguous tensor")
    _chk(children.is_cuda, "children must be a CUDA tensor")
    _chk(children.dtype == torch.int32, "children must be an int tensor")  # the reference reinterprets blindly
    voxelsize, n_max = float(voxelsize), int(n_max)
    b, m = ray_start.shape[0], ray_start.shape[1]
    if shared_tree:
        T, stride, trees = points.shape[-2], 0, 1
    else:
        _chk(points.dim() == 3 and points.shape[0] == b, "points must be [B, T, 3] with B == ray_start.size(0)")
        T, stride, trees = points.shape[1], points.shape[1], b
    _chk(children.shape[-2] == T and children.shape[-1] == 9, "children must be [.., T, 9]")
    idx, dmin, dmax = _hit_outputs(ray_start, n_max)
    with torch.cuda.device(ray_start.device):
        ws_bytes = _L.nsvf_svo_workspace_bytes(T, trees)
        ws = _workspace(ws_bytes, ray_start.device)
        _lib.check(_L.nsvf_svo_intersect(
            _lib.current_stream(ray_start.device), b, T, m, voxelsize, n_max, _p(ray_start), _p(ray_dir),
            _p(points), _p(children), stride, _p(idx), _p(dmin), _p(dmax), _p(ws), ws.numel()))
    return idx, dmin, dmax


class SvoIndex:
    """An octree prepared for intersection (nsvf_svo_prepare): packed nodes, consistency flags, DFS ranks of the leaves
    and their lattice, plus the arrays it was built from.  Build once per octree change, pass as `index=`."""

    def __init__(self, points, children, voxelsize, shared_tree=False):
        _check_float_cuda(points=points)
        _chk(children.is_contiguous() and children.is_cuda and children.dtype == torch.int32,
             "children must be a contiguous CUDA int tensor")
        self.points, self.children, self.voxelsize = points, children, float(voxelsize)
        if shared_tree:
            self.T, self.stride, self.trees = points.shape[-2], 0, 1
        else:
            _chk(points.dim() == 3, "points must be [B, T, 3]")
            self.T, self.stride, self.trees = points.shape[1], points.shape[1], points.shape[0]
        _chk(children.shape[-2] == self.T and children.shape[-1] == 9, "children must be [.., T, 9]")
        with torch.cuda.device(points.device):
            self.ws = _workspace(_L.nsvf_svo_sorted_workspace_bytes(self.T, self.trees, 0), points.device)
            _lib.check(_L.nsvf_svo_prepare(_lib.current_stream(points.device), self.trees, self.T, self.voxelsize,
                                           _p(points), _p(children), self.stride, _p(self.ws), self.ws.numel()))

    def query(self, ray_start, ray_dir, n_max, empty_depth):
        _check_float_cuda(ray_start=ray_start, ray_dir=ray_dir)
        b, m = ray_start.shape[0], ray_start.shape[1]
        _chk(self.stride == 0 or b == self.trees, "one octree per batch row was prepared")
        dev = ray_start.device
        idx, dmin, dmax = _hit_outputs(ray_start, int(n_max))
        hits = torch.empty((b, m), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            scratch = _workspace(_L.nsvf_svo_ray_scratch_bytes(self.trees, b * m), dev)
            _lib.check(_L.nsvf_svo_intersect_sorted_prepared(
                _lib.current_stream(dev), b, self.T, m, self.voxelsize, int(n_max), float(empty_depth), _p(ray_start),
                _p(ray_dir), _p(self.points), _p(self.children), self.stride, _p(idx), _p(dmin), _p(dmax), _p(hits),
                _p(self.ws), self.ws.numel(), _p(scratch), scratch.numel()))
        return idx, dmin, dmax, hits.bool()


def svo_intersect_sorted(ray_start, ray_dir, points, children, voxelsize, n_max, empty_depth=10000.0, shared_tree=False,
                         index=None):
    """Extension: svo_intersect + sort_hits_by_depth (the octree branch of SparseVoxelEncoder.ray_intersect,
    encoder.py:495-524) in one call -> idx i32, min_depth, max_depth f32 [B,M,n_max] sorted by entry depth, hits bool [B,M].
    `index` (SvoIndex of the same octree) skips the per-call preparation."""
    if index is not None:
        return index.query(ray_start, ray_dir, n_max, empty_depth)
    _check_float_cuda(ray_start=ray_start, ray_dir=ray_dir, points=points)
    _chk(children.is_contiguous(), "children must be a contiguous tensor")
    _chk(children.is_cuda, "children must be a CUDA tensor")
    _chk(children.dtype == torch.int32, "children must be an int tensor")
    voxelsize, n_max = float(voxelsize), int(n_max)
    b, m = ray_start.shape[0], ray_start.shape[1]
    if shared_tree:
        T, stride, trees = points.shape[-2], 0, 1
    else:
        _chk(points.dim() == 3 and points.shape[0] == b, "points must be [B, T, 3] with B == ray_start.size(0)")
        T, stride, trees = points.shape[1], points.shape[1], b
    _chk(children.shape[-2] == T and children.shape[-1] == 9, "children must be [.., T, 9]")
    idx, dmin, dmax = _hit_outputs(ray_start, n_max)
    hits = torch.empty((b, m), dtype=torch.uint8, device=ray_start.device)
    with torch.cuda.device(ray_start.device):
        ws = _workspace(_L.nsvf_svo_sorted_workspace_bytes(T, trees, b * m), ray_start.device)
        _lib.check(_L.nsvf_svo_intersect_sorted(
            _lib.current_stream(ray_start.device), b, T, m, voxelsize, n_max, float(empty_depth), _p(ray_start),
            _p(ray_dir), _p(points), _p(children), stride, _p(idx), _p(dmin), _p(dmax), _p(hits), _p(ws), ws.numel()))
    return idx, dmin, dmax, hits.bool()


def ball_intersect(ray_start, ray_dir, points, radius, n_max):
    """fairnr/clib/src/intersect.cpp:15-44 (no caller in the reference; provided for API completeness)."""
    _check_float_cuda(ray_start=ray_start, ray_dir=ray_dir, points=points)
    radius, n_max = float(radius), int(n_max)
    b, m = ray_start.shape[0], ray_start.shape[1]
    _chk(points.dim() == 3 and points.shape[0] == b, "points must be [B, n, 3] with B == ray_start.size(0)")
    n = points.shape[1]
    idx, dmin, dmax = _hit_outputs(ray_start, n_max)
    with torch.cuda.device(ray_start.device):
        _lib.check(_L.nsvf_ball_intersect(_lib.current_stream(ray_start.device), b, n, m, radius, n_max, _p(ray_start),
                                          _p(ray_dir), _p(points), n * 3, _p(idx), _p(dmin), _p(dmax)))
    return idx, dmin, dmax


def triangle_intersect(ray_start, ray_dir, face_points, cagesize, blur, n_max):
    """fairnr/clib/src/intersect.cpp:120-146.  face_points f32 [B, n, 9] -> idx i32 [B,M,n_max],
    depth f32 [B,M,n_max*3], uv f32 [B,M,n_max*2]."""
    _check_float_cuda(ray_start=ray_start, ray_dir=ray_dir, face_points=face_points)
    cagesize, blur, n_max = float(cagesize), float(blur), int(n_max)
    b, m = ray_start.shape[0], ray_start.shape[1]
    _chk(face_points.dim() == 3 and face_points.shape[0] == b, "face_points must be [B, n, 9]")
    n = face_points.shape[1]
    dev = ray_start.device
    idx = torch.empty((b, m, n_max), dtype=torch.int32, device=dev)
    depth = torch.empty((b, m, n_max * 3), dtype=torch.float32, device=dev)
    uv = torch.empty((b, m, n_max * 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_L.nsvf_triangle_intersect(_lib.current_stream(dev), b, n, m, cagesize, blur, n_max, _p(ray_start),
                                              _p(ray_dir), _p(face_points), n * 9, _p(idx), _p(depth), _p(uv)))
    return idx, depth, uv


def uniform_ray_sampling(pts_idx, min_depth, max_depth, uniform_noise, step_size, max_steps):
    """fairnr/clib/src/sample.cpp:23-55.  [G,R,P] inputs, noise [G,R,max_steps] -> 3 x [G,R,max_steps]."""
    _check(dict(min_depth=min_depth, max_depth=max_depth, uniform_noise=uniform_noise), dict(pts_idx=pts_idx))
    step_size, max_steps = float(step_size), int(max_steps)
    g, r, p = min_depth.shape
    dev = pts_idx.device
    sidx = torch.empty((g, r, max_steps), dtype=torch.int32, device=dev)
    sdepth = torch.empty((g, r, max_steps), dtype=torch.float32, device=dev)
    sdists = torch.empty((g, r, max_steps), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_L.nsvf_uniform_ray_sampling(
            _lib.current_stream(dev), g, r, p, max_steps, step_size, _p(pts_idx), _p(min_depth), _p(max_depth),
            _p(uniform_noise), _p(sidx), _p(sdepth), _p(sdists), None))
    return sidx, sdepth, sdists


def inverse_cdf_sampling(pts_idx, min_depth, max_depth, uniform_noise, probs, steps, fixed_step_size):
    """fairnr/clib/src/sample.cpp:58-95.  max_steps = uniform_noise.size(-1)."""
    _check(dict(min_depth=min_depth, max_depth=max_depth, uniform_noise=uniform_noise, probs=probs, steps=steps),
           dict(pts_idx=pts_idx))
    fixed_step_size = float(fixed_step_size)
    g, r, p = min_depth.shape
    max_steps = uniform_noise.shape[-1]
    dev = pts_idx.device
    sidx = torch.empty((g, r, max_steps), dtype=torch.int32, device=dev)
    sdepth = torch.empty((g, r, max_steps), dtype=torch.float32, device=dev)
    sdists = torch.empty((g, r, max_steps), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_L.nsvf_inverse_cdf_sampling(
            _lib.current_stream(dev), g, r, -1, 0, p, max_steps, fixed_step_size, _p(pts_idx), _p(min_depth),
            _p(max_depth), _p(uniform_noise), 0.5, _p(probs), _p(steps), _p(sidx), _p(sdepth), _p(sdists), None))
    return sidx, sdepth, sdists


def build_octree(center, points, depth):
    """fairnr/clib/src/octree.cpp:125-135.  center [3] (any real dtype), points i64 [n,3], depth int ->
    (centers i32 [T,3], children i32 [T,9]) on center's device.  The tree is built on the host from one
    D2H copy (the reference syncs once per point per level through .item())."""
    depth = int(depth)
    dev = center.device
    c = center.detach().to("cpu", torch.float64).contiguous()
    p = points.detach().to("cpu", torch.int64).contiguous()
    _chk(p.dim() == 2 and p.shape[1] == 3, "points must be [n, 3]")
    import ctypes
    total, term = ctypes.c_longlong(0), ctypes.c_longlong(0)
    _lib.check(_L.nsvf_octree_build(c.data_ptr(), p.data_ptr(), p.shape[0], depth,
                                    ctypes.addressof(total), ctypes.addressof(term)))
    T = total.value
    centers = torch.empty((T, 3), dtype=torch.int32)
    children = torch.empty((T, 9), dtype=torch.int32)
    _lib.check(_L.nsvf_octree_flatten(centers.data_ptr(), children.data_ptr(), T))
    return centers.to(dev), children.to(dev)
