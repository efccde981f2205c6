"""Half-voxel splitting and pruning on the sm_100a kernels of csrc/split_prune.cu (host-side glue)."""
import ctypes

import torch

from . import _lib

_L = _lib.load()
_p = _lib.ptr


def splitting_points(point_xyz, point_feats, values, half_voxel):
    """See geometry.splitting_points.  point_xyz f32 [n,3] (CUDA), point_feats int [n,8], values f32 [Kc,D] or None."""
    if not point_xyz.is_cuda:
        raise RuntimeError("nsvf_b200: splitting_points needs CUDA tensors (there is no CPU path)")
    dev = point_xyz.device
    pts = point_xyz.detach().float().contiguous()
    half_voxel = float(half_voxel)
    quarter = torch.tensor(half_voxel, dtype=torch.float32).mul(0.5).item()       # float32 product, like the reference
    n = pts.size(0)
    # host-side extent of the key lattice (2 tiny reductions + one sync; splitting runs 3 times per training)
    pmin_t = pts.min(dim=0)[0]
    max_coord_t = ((pts - pmin_t) / quarter).round_().max(dim=0)[0]
    pmin = (ctypes.c_float * 3)(*pmin_t.tolist())
    max_coord = (ctypes.c_int * 3)(*[int(v) for v in max_coord_t.tolist()])
    with torch.cuda.device(dev):
        st = _lib.current_stream(dev)
        ws = torch.empty(_L.nsvf_split_workspace_bytes(max_coord), dtype=torch.uint8, device=dev)
        new_points = torch.empty((8 * n, 3), dtype=torch.float32, device=dev)
        n_keys_t = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(_L.nsvf_split_mark(st, n, _p(pts), half_voxel, pmin, max_coord, _p(new_points), _p(n_keys_t),
                                      _p(ws), ws.numel()))
        n_keys = int(n_keys_t.item())
        new_feats = torch.empty((8 * n, 8), dtype=torch.int32, device=dev)
        parent = torch.empty(n_keys, dtype=torch.int32, device=dev)
        new_keys = torch.empty((n_keys, 3), dtype=torch.int32, device=dev)
        new_values, feats32, vals, D = None, None, None, 0
        if values is not None:
            vals = values.detach().float().contiguous()
            D = vals.size(-1)
            feats32 = point_feats.to(torch.int32).contiguous()
            new_values = torch.empty((n_keys, D), dtype=torch.float32, device=dev)
        _lib.check(_L.nsvf_split_emit(st, n, D, _p(pts), _p(feats32), _p(vals), half_voxel, pmin, max_coord, n_keys,
                                      _p(new_feats), _p(parent), _p(new_keys), _p(new_values), _p(ws), ws.numel()))
    return new_points.type_as(point_xyz), new_feats.long(), new_values, new_keys.long()


def lattice_embed(feats32, centres, values, voxel_size, v0, nv, bits):
    """emb f32 [nv * bits^3, D] at the lattice points of voxels [v0, v0+nv) (encoder.get_scores' field inputs)."""
    dev = values.device
    D = values.size(-1)
    out = torch.empty((nv * bits ** 3, D), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_L.nsvf_prune_lattice_embed(_lib.current_stream(dev), nv, v0, bits, D, _p(feats32), _p(centres),
                                               _p(values), float(voxel_size), _p(out)))
    return out


def prune_keep(sigma, n_points, th):
    """(keep bool [nv], min_score f32 [nv]) from sigma f32 [nv * n_points]."""
    sigma = sigma.detach().float().contiguous()
    nv = sigma.numel() // n_points
    dev = sigma.device
    keep = torch.empty(nv, dtype=torch.uint8, device=dev)
    score = torch.empty(nv, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_L.nsvf_prune_keep(_lib.current_stream(dev), nv, n_points, _p(sigma), float(th), _p(keep), _p(score)))
    return keep.bool(), score
