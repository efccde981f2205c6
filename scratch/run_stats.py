"""Structure of real ray-marched sample streams (CPU oracle): run length per voxel and corner sharing between consecutive
runs of a ray — the quantities that bound the reduction traffic of trilinear_bwd."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import oracle
from oracle import wrappers
from nsvf_b200 import synthetic
oracle.build()
for name, nrays in (("C3", 3000), ("C2", 3000)):
    scene = synthetic.make_scene(name)
    pts = scene.points.copy(); pts[:, 0] += np.float32(scene.voxel_size / 10)
    rs, rd = synthetic.camera_rays(800, 800, 1, radius=4.5 if name == "C3" else 3.2, seed=3)
    sel = np.random.RandomState(0).choice(800 * 800, nrays, replace=False)
    o = np.broadcast_to(rs.numpy(), (1, 800 * 800, 3))[:, sel].reshape(1, -1, 3).astype(np.float32)
    d = rd.numpy()[:, sel].reshape(1, -1, 3).astype(np.float32)
    idx, dmin, dmax = oracle.aabb_intersect(o, d, pts, scene.voxel_size, scene.max_hits)
    idx_t, dmin_t, dmax_t, hits = wrappers.sort_hits(*[torch.from_numpy(a[0]) for a in (idx, dmin, dmax)])
    idx_t, dmin_t, dmax_t = idx_t[hits], dmin_t[hits], dmax_t[hits]
    probs, stp = wrappers.probs_and_steps(idx_t, dmin_t, dmax_t, scene.step_size)
    sidx, sdep, sdist = wrappers.inverse_cdf_sampling(wrappers.NumpyExt(), idx_t, dmin_t, dmax_t, probs, stp, -1, True)
    sidx = sidx.numpy()
    feats = scene.feats
    runs, shared, nsamp = [], [], 0
    for r in range(sidx.shape[0]):
        v = sidx[r][sidx[r] >= 0]
        if len(v) == 0: continue
        nsamp += len(v)
        starts = np.flatnonzero(np.r_[True, v[1:] != v[:-1]])
        lens = np.diff(np.r_[starts, len(v)])
        runs += lens.tolist()
        vv = v[starts]
        for a, b in zip(vv[:-1], vv[1:]):
            shared.append(len(set(feats[a].tolist()) & set(feats[b].tolist())))
    runs, shared = np.array(runs), np.array(shared)
    flush_rows_now = 8 * len(runs)
    flush_rows_merged = 8 * len(runs) - shared.sum()
    print("%s: %d rays hit, %d samples, %d voxel runs; run length mean %.2f median %d p10 %d p90 %d; "
          "consecutive runs of a ray share %.2f corners on average (>=4: %.0f %%); reduction rows per sample: %.2f now "
          "(one flush of 8 rows per run) -> %.2f with shared corners merged (%.0f %% fewer)" %
          (name, sidx.shape[0], nsamp, len(runs), runs.mean(), np.median(runs), np.percentile(runs, 10), np.percentile(runs, 90),
           shared.mean(), 100 * (shared >= 4).mean(), flush_rows_now / nsamp, flush_rows_merged / nsamp,
           100 * (1 - flush_rows_merged / flush_rows_now)))
