#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout (rank 0).

  python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port, host cores)

Workload (config.workload = "C2"): the `nsvf_base` training step of BASELINE.json configs[1] —
Synthetic-NeRF-shaped scene (bbox centres +-1.2, voxel 0.4 -> 343 voxels, step = voxel/8, max_hits 60),
4 views x 800x800 rays per GPU intersected (`--no-sampling-at-reader`), 4 x 2048 rays sampled from the hit mask
and marched (inverse-CDF sampling -> trilinear interpolation -> field MLP -> compositing), loss, backward, Adam.
Random-init weights, synthetic cameras / targets (no datasets offline).  The field MLP (545 297 parameters,
fp32): its contractions run on torch/cuBLAS tensor cores as BASELINE.json prescribes (cuBLAS 12.9's fp32-accurate
BF16x9 algorithm, nsvf_b200/blas.py; --mlp-gemm simt for the plain SGEMM), the LayerNorm/ReLU passes around them
and everything else are hand-written kernels.

  value  = marched rays per second, whole job (all ranks), inputs resident in HBM
  e2e    = same, through the public pipeline call with HOST (pinned) rays + targets copied in every step and the
           loss read back
  roofline = the hand-written kernel with the largest share of the step's device time (ln_relu_bwd_kernel, the fused
           LayerNorm+ReLU backward around the MLP's cuBLAS GEMMs), timed live with CUDA events recorded around the
           launch inside the timed steps (nsvf_profile_kernel hook); roofline_hot_path = the same for the dominant
           kernel of the ray-marching path itself (the any-hit intersection); roofline_at_scale = the HBM-bound
           kernels of the path at full-frame sizes
  cpu_baseline = the oracle port (oracle/, C + OpenMP + torch-CPU MLP) on a bounded sample of the same step
  frame  = ms per 800x800 frame on the C3 scene (~112k voxels, eval, early termination 0.01), extra key
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# cuBLAS for the field MLP's fp32 contractions: must be chosen before torch is imported (nsvf_b200/blas.py).
# --mlp-gemm simt keeps torch's bundled cuBLAS and plain SGEMM; the default runs the same fp32 GEMMs on the BF16
# tensor cores with cuBLAS 12.9's fp32-accurate BF16x9 algorithm.  The CPU reference arm is not affected.
from nsvf_b200 import blas as _blas
if "reference" not in sys.argv[1:] and "simt" not in sys.argv[1:] and "--mlp-gemm=simt" not in sys.argv[1:]:
    _blas.use_system_cublas()

import numpy as np
import torch

VIEWS, RES, PIX_PER_VIEW = 4, 800, 2048
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the frame's kernels, `ncu --set full` on the C3 hot-path frame
# (profiles/r2_ncu_c3_frame.md and r2_ncu_c3_frame_final.md; same scene, camera and chunking as frame_bench below)
NCU_TRAFFIC = {"inverse_cdf_sampling_kernel": 7.305e9, "march_composite_fwd_kernel": 0.983e9,
               "march_compact_kernel": 21.8e6, "trilinear_fwd_kernel": 18.7e6, "march_epilogue_kernel": 14.7e6,
               "aabb_intersect_sorted_kernel": 2.139e9, "march_transpose_kernel": None, "inverse_cdf_plan_kernel": 0.840e9,
               "inverse_cdf_stream_kernel": 0.165e9}   # stream: 2.31 GB over the frame's 14 launches (the first one 1.83 GB)
LN_BWD_DRAM_TRAFFIC = 149.0e6  # bytes per launch: dram read 135 MB + write 14 MB, ncu --set full, [65536, 256] (profiles/r1b_ncu_ln_relu.md)
METRIC = "rays/s (intersect+sample+composite), nsvf_base training step"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C5"],
                    help="BASELINE.json configs[1..4]; C2 (the nsvf_base training step) is the headline the driver runs, "
                         "the others are the multi-GPU runs of configs[2..4] (frames sharded / octree + sharded pruning / "
                         "hierarchical training)")
    ap.add_argument("--no-frame", action="store_true", help="skip the C3 full-frame extra")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-on-this-GPU legs (child process)")
    ap.add_argument("--no-stages", action="store_true", help="skip the informational per-stage timings")
    ap.add_argument("--mlp-gemm", default="bf16x9", choices=["bf16x9", "simt"],
                    help="fp32 GEMMs of the field MLP: cuBLAS 12.9 BF16x9 emulation on tensor cores (fp32-accurate) "
                         "or torch's bundled cuBLAS SGEMM on the SIMT pipe")
    ap.add_argument("--no-graphs", action="store_true", help="launch the field MLP kernel by kernel instead of replaying "
                                                               "one CUDA graph per renderer chunk")
    ap.add_argument("--cpu-fraction", type=int, default=16, help="CPU arms run 1/FRACTION of the rays per step")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region through NVML (same counters nvidia-smi prints)
# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def reset(self):
        self.samples, self.reasons = [], set()

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# allocator pre-sizing
# ------------------------------------------------------------------------------------------------------
def presize_allocator(dev, large_gib=12, small_mib=1024):
    """Fill torch's caching allocator BEFORE the NCCL communicator exists, so that no step ever calls cudaMalloc.

    Measured on 2-GPU boxes (profiles/r2_multirank_stall.md): once several processes share the box with peer access
    enabled, ONE cudaMalloc takes 40 - 100 ms (1 ms in a single process), and it is device-synchronising — a step that
    needs a new 2 MiB small-pool segment stalls, and through the gradient all-reduce every rank stalls with it.  That
    is the multi-rank "stall" of round 1 (0.2 - 12 s at 8 ranks).  The large pool is served by one big block that the
    allocator splits on demand; the small pool (tensors < 1 MiB: ray offsets, per-ray state, scalars) cannot use that
    block, so it gets its own segments."""
    big = torch.empty(large_gib << 30, dtype=torch.uint8, device=dev)
    small = [torch.empty(512 << 10, dtype=torch.uint8, device=dev) for _ in range(small_mib * 2)]
    del big, small


# ------------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------------
def make_batches(n_batches, rank, pinned):
    from nsvf_b200 import synthetic
    out = []
    for b in range(n_batches):
        rs, rd = synthetic.camera_rays(RES, RES, VIEWS, radius=3.2, seed=137 * b + rank)   # unique_seed = step*137+rank
        g = torch.Generator().manual_seed(1000 + 137 * b + rank)
        target = torch.rand(VIEWS * RES * RES, 3, generator=g) * 2 - 1                    # min_color = -1 range
        rs, rd = rs[None, :, None, 0, :].contiguous(), rd[None].contiguous()               # [1,V,1,3], [1,V,P,3]
        if pinned:
            rs, rd, target = rs.pin_memory(), rd.pin_memory(), target.pin_memory()
        out.append((rs, rd, target))
    return out


def build_model(device, scene_name="C2", train=True, field="mlp", tolerance=0.0, chunk=64, sigma_bias=0.0, seed=0,
                graphs=False):
    from nsvf_b200 import synthetic
    from nsvf_b200.encoder import SparseVoxelEncoder
    from nsvf_b200.field import RadianceField, TrivialField
    from nsvf_b200.renderer import VolumeRenderer
    from nsvf_b200.pipeline import NSVFPipeline
    torch.manual_seed(seed)
    scene = synthetic.make_scene(scene_name)
    enc = SparseVoxelEncoder(scene.points, scene.voxel_size, max_hits=scene.max_hits)
    fld = RadianceField(sigma_bias=sigma_bias) if field == "mlp" else TrivialField()
    if graphs and field == "mlp" and train:
        from nsvf_b200.field import GraphedField
        fld = GraphedField(fld, rows=1024 * chunk, eager_first=1)   # first chunk eager: live kernel timing
    ren = VolumeRenderer(chunk_size=chunk, discrete_regularization=train, raymarching_tolerance=tolerance)
    pipe = NSVFPipeline(enc, fld, ren, pixel_per_view=PIX_PER_VIEW if train else 0).to(device)
    return pipe.train(train), scene


def loss_fn(out, target):
    sel = target if out["sampled"] is None else target[out["sampled"].reshape(-1)]
    rgb = ((out["colors"] - sel) ** 2).mean() * 128.0          # --color-weight 128
    alpha = (out["missed"] ** 2).mean()                        # --alpha-weight 1 (all sampled rays hit the object mask)
    return rgb + alpha


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from nsvf_b200 import _lib
    L = _lib.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pipe, scene = build_model(dev, graphs=not args.no_graphs)
    if hasattr(pipe.field, "capture"):
        pipe.field.capture(dev)          # CUDA graphs of the field's chunk forward / backward, before any eager pass
    # order matters: graph capture empties the allocator's cache (torch.cuda.graph.__enter__ calls empty_cache), and a
    # cudaMalloc costs 40 - 100 ms once the NCCL communicator has mapped the peers — so: capture, pre-size, then NCCL
    t_pre = time.perf_counter()
    presize_allocator(dev)
    torch.cuda.synchronize()
    t_pre = time.perf_counter() - t_pre
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if os.environ.get("NSVF_BENCH_TRACE"):
        print("rank %d: allocator pre-sized in %.2f s" % (rank, t_pre), file=sys.stderr, flush=True)
    model = pipe
    params = [p for p in pipe.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3, betas=(0.9, 0.999))
    # gradient exchange: ONE flat fp32 bucket holding values.weight.grad + the MLP grads (2.3 MB), every
    # parameter's .grad a view into it, one NCCL all-reduce (AVG) per step right after backward.  No DDP: its reducer
    # hooks / bucket rebuild bought nothing for a 2.3 MB payload and sat on the per-rank host path.
    flat_grad = torch.zeros(sum(p.numel() for p in params), device=dev)
    off = 0
    for p in params:
        p.grad = flat_grad[off: off + p.numel()].view_as(p)
        off += p.numel()
    host = make_batches(2, rank, pinned=True)
    resident = [tuple(t.to(dev) for t in b) for b in host]
    staging = tuple(torch.empty_like(t, device=dev) for t in host[0])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step(rs, rd, target):
        flush.zero_()
        out = model(rs, rd)
        loss = loss_fn(out, target)
        flat_grad.zero_()                 # grads are views of the flat bucket: backward accumulates into it
        loss.backward()
        if world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG)
        opt.step()
        return loss, out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(n, from_host, kname=b"ln_relu_bwd_kernel"):
        """n steps; returns (ms total via CUDA events on the launching stream, list of the ms of kernel `kname`:
        its last launch of each step, bracketed by events recorded on the launching stream by the library)."""
        k_evs = []
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host0 = time.perf_counter()
        trace = os.environ.get("NSVF_BENCH_TRACE")
        if trace:
            import faulthandler
            tb = open(os.path.join(ROOT, "gpurun_out", "bench_stall_rank%d.txt" % rank), "a")
        t_prev = t_host0
        for i in range(n):
            if trace:      # stall hunt: dump the Python stack of a step that takes > 90 ms, and count cudaMallocs
                faulthandler.dump_traceback_later(0.09, exit=False, file=tb)
                m0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); b.record()                                    # materialise the cudaEvent_t handles
            _lib.check(L.nsvf_profile_kernel(kname, a.cuda_event, b.cuda_event))
            if from_host:
                hb = host[i % len(host)]
                for d, s in zip(staging, hb):
                    d.copy_(s, non_blocking=True)                     # H2D from pinned memory, inside the timed region
                loss, out = step(*staging)
                loss.item()                                           # D2H read of the step's result
            else:
                loss, out = step(*resident[i % len(resident)])
            k_evs.append((a, b))
            if trace:
                faulthandler.cancel_dump_traceback_later()
                now = time.perf_counter()
                ms_ = torch.cuda.memory_stats(dev)
                print("rank %d step %d from_host=%s: host %.2f ms (+%d cudaMalloc) segs small %d large %d reserved %d MiB "
                      "peak-allocated %d MiB" %
                      (rank, i, from_host, (now - t_prev) * 1e3, ms_.get("num_device_alloc", 0) - m0,
                       ms_.get("segment.small_pool.current", 0), ms_.get("segment.large_pool.current", 0),
                       ms_.get("reserved_bytes.all.current", 0) >> 20, ms_.get("allocated_bytes.all.peak", 0) >> 20),
                      file=sys.stderr, flush=True)
                t_prev = now
        e1.record()
        host_ms[0] = (time.perf_counter() - t_host0) * 1e3 / n      # host time to ENQUEUE a step (no sync)
        barrier()
        L.nsvf_profile_kernel(None, None, None)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, [a.elapsed_time(b) for a, b in k_evs], float(loss.item()), out

    # shape warm-up: the eager chunks of a step (the first one, padded to 65536 rows, and a short last one whose row count
    # changes every step) can make cuBLAS pick a GEMM kernel it has not used yet, and CUDA loads kernels lazily — a
    # first use in the middle of the timed steps costs tens of ms (more with several processes on the box).  One eager
    # forward + backward of the field per row-count bucket, via autograd.grad so that DDP's hooks stay out of it.
    try:
        inner = getattr(pipe.field, "field", pipe.field)
        params = [p for p in inner.parameters() if p.requires_grad]
        for M in list(range(2048, 65536 + 1, 2048)):
            emb = torch.zeros(M, 32, device=dev, requires_grad=True)
            ray = torch.nn.functional.normalize(torch.ones(M, 3, device=dev), dim=-1)
            o = inner({"emb": emb, "ray": ray})
            torch.autograd.grad(o["sigma"].sum() + o["texture"].sum(), params + [emb])
        del emb, ray, o
        torch.cuda.synchronize()
    except Exception as e:      # warm-up only: never fatal
        print("shape warm-up skipped: %r" % (e,), file=sys.stderr)
    import gc
    gc.collect()
    gc.freeze()          # everything built so far is permanent: keeps full collections out of the timed steps
    # the clock sampler starts before the warm-up steps: the first NVML queries of a process are slow and contend with
    # the driver (they cost the first timed pass ~50 ms on a fresh box); its samples are reset when the timed steps start
    sampler = ClockSampler(local)
    sampler.start()
    n_warm = max(args.warmup, 3)
    timed(n_warm, False)
    sampler.reset()
    launches0 = L.nsvf_kernel_launches()
    replays0 = getattr(pipe.field, "graph_replays", 0)
    ms, kms, loss_val, out = timed(args.steps, False)
    # kernels of libnsvf_b200.so launched in the timed region: from the host + inside the replayed chunk graphs
    replays = getattr(pipe.field, "graph_replays", 0) - replays0
    launches = L.nsvf_kernel_launches() - launches0 + replays * getattr(pipe.field, "launches_per_replay", 0)
    host_enqueue_ms = host_ms[0]
    clocks = sampler.stop()
    from nsvf_b200 import ops as _ops
    ln_M, ln_N = _ops.LAST_LN_BWD_SHAPE      # the launch the events bracketed: first layer of the step's first chunk
    timed(max(args.warmup, 3), True, b"aabb_hit_mask_kernel")      # the end-to-end pass gets its own W warm-up steps
    ms_e2e, kms_hit, _, _ = timed(args.steps, True, b"aabb_hit_mask_kernel")

    rays_marched = VIEWS * PIX_PER_VIEW
    rays_intersected = VIEWS * RES * RES
    value = world * rays_marched * args.steps / (ms / 1e3)
    e2e_value = world * rays_marched * args.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    # roofline of the dominant hand-written kernel of the step = the one with the largest share of device time among
    # ours (profiles/r1b_launches_step.md): ln_relu_bwd_kernel<256>, 63 launches per step.  Algorithmic bytes per
    # launch: read h and dy, write dh (3 x 4 B per element) + mean/rstd (8 B per row) + the per-CTA
    # partial column sums.  Timed live: the last launch of every step (M = ln_M rows, the first layer of the first
    # chunk) between two CUDA events the library records on the launching stream.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    step_ms = ms / args.steps
    ln_ws = L.nsvf_ln_relu_bwd_workspace_bytes(ln_M, ln_N)
    alg_bytes = 12 * ln_M * ln_N + 8 * ln_M + ln_ws + 8 * ln_N
    k_ms = float(np.mean(kms))
    achieved = alg_bytes / (k_ms / 1e3) / 1e9
    # launches of this kernel per step: one per LayerNorm'ed FCLayer of the field and renderer chunk (the chunks replay
    # CUDA graphs, so the library's per-launch hook cannot count them)
    n_fc = sum(1 for m in pipe.field.modules() if type(m).__name__ == "_FCLayer")
    ln_launches = n_fc * max(1, -(-int(out["ae"]) // 65536))
    roofline = {"kernel": "ln_relu_bwd_kernel<%d> (LayerNorm+ReLU backward of an FCLayer, [%d, %d] fp32)" % (ln_N, ln_M, ln_N),
                "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4),
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape, one `ncu --set full`
                # capture (profiles/r1b_ncu_ln_relu.md), per launch
                "traffic": LN_BWD_DRAM_TRAFFIC, "peak_source": peak_src,
                "note": "dy was written by the preceding cuBLAS GEMM and is partly still in the 126 MB L2, so DRAM "
                        "traffic is below the algorithmic bytes and `achieved` can exceed what DRAM alone delivers",
                "kernel_ms": round(k_ms, 4), "algorithmic_bytes_per_launch": alg_bytes,
                "launches_per_step": ln_launches, "share_of_step": round(ln_launches * k_ms / step_ms, 4)}
    # the dominant kernel of the ray-marching path proper: the any-hit intersection over all V*H*W rays,
    # 24 B/ray in + 1 B/ray out (+ 12 B/voxel once).  ALU/issue-bound by design (intersection is not
    # HBM-bound); the HBM-bound kernels of the path are measured at full-frame sizes under "roofline_at_scale".
    P = scene.max_hits
    hit_bytes = rays_intersected * (24 + 1) + 12 * scene.n
    hit_ms = float(np.mean(kms_hit))
    roofline_hot = {"kernel": "aabb_hit_mask_kernel (any-hit intersection of all %d rays)" % rays_intersected,
                    "bound": "hbm", "achieved": round(hit_bytes / (hit_ms / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(hit_bytes / (hit_ms / 1e3) / 1e9 / peak, 4), "traffic": 61.9e6,
                    "note": "ALU/issue-bound by design (73 % issue-active in ncu, DRAM traffic == algorithmic bytes)",
                    "kernel_ms": round(hit_ms, 4), "algorithmic_bytes_per_launch": hit_bytes,
                    "share_of_step": round(hit_ms / step_ms, 4)}

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: nsvf_base training step, 343 voxels (voxel 0.4, step 1/8, max_hits 60), "
                               "4 views x 800x800 rays intersected + 4 x 2048 rays marched per GPU, fwd+bwd+Adam, "
                               "field MLP fp32 on cuBLAS", "mlp_gemm": _blas.mode(),
                   "mlp_launch": ("one CUDA graph per full renderer chunk (forward + backward), first chunk eager"
                                  if hasattr(pipe.field, "graph_replays") else "kernel by kernel"),
                   "rays_intersected_per_step_per_gpu": rays_intersected,
                   "rays_marched_per_step_per_gpu": rays_marched, "samples_evaluated_per_step": int(out["ae"]),
                   "l2": "256 MiB memset at the start of every step (inside the timed region)",
                   "parallelism": "dp%d (rays sharded by view, voxel set replicated, NCCL grad all-reduce)" % world},
        "clocks": clocks, "gpu_launches": int(launches), "cuda_graph_replays": int(replays),
        "host_enqueue_ms_per_step": round(host_enqueue_ms, 3),
        "e2e": {"value": round(e2e_value, 1), "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 4)},
        "roofline": roofline, "roofline_hot_path": roofline_hot, "loss": round(loss_val, 5),
    }

    if rank == 0 and not args.no_stages:
        try:
            line["roofline_at_scale"] = roofline_at_scale(dev, peak)
        except Exception as e:
            line["roofline_at_scale"] = {"error": repr(e)[:200]}
        try:
            line["stages_ms"] = stage_times(dev, pipe, resident[0])
        except Exception as e:   # informational only
            line["stages_ms"] = {"error": repr(e)[:200]}
    if world > 1:
        dist.barrier()
    # the ray-marching path alone ("MLP ... once excluded with a trivial field_fn"): the same step with a
    # contraction-free stand-in field, so that intersect + sample + interpolate + composite (fwd + bwd) + Adam on the
    # voxel embeddings is all that is timed
    try:
        hot = hot_path_step(dev, resident, args.steps, max(args.warmup, 3))
        line["value_hot_path"] = {"value": round(world * rays_marched / (hot["ms_per_step"] / 1e3), 1), "unit": "rays/s",
                                  "ms_per_step": hot["ms_per_step"], "field": "trivial (no contraction)",
                                  "per_gpu": True if world > 1 else False}
    except Exception as e:
        hot = None
        line["value_hot_path"] = {"error": repr(e)[:200]}
    if rank == 0 and not args.no_frame:
        # BASELINE.json configs[2]: one 800x800 frame of the C3 scene (eval, early termination 0.01, chunk 512), with the
        # field MLP and with the stand-in field; the §8 kernels are timed live inside the hot-path frame
        c3 = None
        try:
            from nsvf_b200 import synthetic
            rs3, rd3 = synthetic.camera_rays(RES, RES, 1, radius=4.5, seed=7, device=dev)
            rs3, rd3 = rs3[None, :, None, 0, :].contiguous(), rd3[None].contiguous()
            pipe_c3, scene3 = build_model(dev, "C3", train=False, field="mlp", tolerance=0.01, chunk=512, sigma_bias=2.0)
            fr = frame_bench(dev, pipe_c3, rs3, rd3)
            c3 = {"pipe": pipe_c3, "scene": scene3, "rs": rs3, "rd": rd3, "frame": fr}
            line["frame"] = {k: v for k, v in fr.items() if not k.startswith("_")}
            kernels = frame_rooflines(dev, pipe_c3, rs3, rd3, peak, peak_src,
                                      fr["hot_path_only(trivial field)"]["ms_per_800x800_frame"])
            for k in kernels:
                k["traffic"] = NCU_TRAFFIC.get(k["kernel"])
            hbm = [k for k in kernels if not k["kernel"].startswith("aabb_intersect")]
            if hbm:
                # THE roofline of this line: the HBM-bound kernel of the ray-marching path with the largest share of the
                # frame's device time; the field-MLP glue kernel that led round 1's line moves to roofline_mlp_glue
                line["roofline_mlp_glue"] = line["roofline"]
                line["roofline"] = max(hbm, key=lambda k: k["share_of_frame"])
            line["roofline_kernels"] = kernels
        except Exception as e:
            line["frame"] = {"error": repr(e)[:300]}
        if world == 1 and not args.no_ref_gpu and c3 is not None:
            try:
                ours = {"step": {"ms_per_step": round(ms / args.steps, 4)},
                        "step_hot_path": {"ms_per_step": hot["ms_per_step"] if hot else None}}
                ref, vs = ref_gpu_leg(dev, pipe, [tuple(t.cpu() for t in b) for b in host], out["sampled"].reshape(-1),
                                      ours, steps=min(args.steps, 6), c3=c3)
                line["ref_gpu"], line["vs_ref_gpu"] = ref, vs
            except Exception as e:
                line["ref_gpu"] = {"error": repr(e)[:300]}
        del c3
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            # bounded sample: ~10-20 s of host-core work (8 steps of 1/4 of the rays each)
            line["cpu_baseline"] = cpu_arm(steps=8, warmup=1, fraction=max(args.cpu_fraction // 4, 1))["cpu_baseline"]
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)[:200]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def hot_path_step(dev, resident, steps, warmup):
    """The C2 training step with the stand-in field (no MLP): device time per step over `steps` steps."""
    pipe, _ = build_model(dev, field="trivial")
    params = [p for p in pipe.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3, betas=(0.9, 0.999))

    def step(i):
        rs, rd, target = resident[i % len(resident)]
        out = pipe(rs, rd)
        loss = loss_fn(out, target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return out
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = step(i)
    e1.record()
    torch.cuda.synchronize()
    return {"ms_per_step": round(e0.elapsed_time(e1) / steps, 4), "samples_evaluated": int(out["ae"])}


def _time(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4)


def real_sample_stream(dev, scene_name="C3", columns=64):
    """Ray-marched samples of an 800x800 view of the scene in ray-major order (what a training step's wide windows hand
    to the interpolation): intersect -> inverse-CDF sampling -> compaction of the first `columns` sample columns."""
    from nsvf_b200 import synthetic, _lib
    from nsvf_b200.encoder import SparseVoxelEncoder
    L, p = _lib.load(), _lib.ptr
    scene = synthetic.make_scene(scene_name)
    enc = SparseVoxelEncoder(scene.points, scene.voxel_size, max_hits=scene.max_hits).to(dev).eval()
    rs, rd = synthetic.camera_rays(RES, RES, 1, radius=4.5, seed=7, device=dev)
    rs, rd = rs[None, :, None, 0, :].contiguous(), rd[None].contiguous()
    with torch.no_grad():
        st = enc.precompute(id=torch.zeros(1, dtype=torch.long, device=dev))
        rs_f, rd_f, inter, hits = enc.ray_intersect(rs, rd, st)
        sel = hits.reshape(-1).nonzero(as_tuple=True)[0]
        inter = {k: v.reshape(-1, v.size(-1)).index_select(0, sel) for k, v in inter.items()}
        dists = (inter["max_depth"] - inter["min_depth"]).masked_fill(inter["intersected_voxel_idx"].eq(-1), 0)
        inter["probs"] = dists / dists.sum(-1, keepdim=True)
        inter["steps"] = dists.sum(-1) / enc.step_size
        smp = enc.ray_sample(inter, trimmed=True)
        sidx, sdep, lens = smp["sampled_point_voxel_idx"], smp["sampled_point_depth"], smp["sampled_point_count"]
        K = min(columns, sidx.shape[1])
        mask = torch.arange(K, device=dev)[None] < lens[:, None]
        o = rs_f.reshape(-1, 3).index_select(0, sel)
        d = rd_f.reshape(-1, 3).index_select(0, sel)
        vox = sidx[:, :K][mask].contiguous()
        xyz = (o[:, None] + d[:, None] * sdep[:, :K, None])[mask].contiguous()
    feats = enc._kept_geometry()[1]
    pts = st["voxel_center_xyz"].reshape(-1, 3).contiguous()
    values = st["voxel_vertex_emb"].reshape(-1, 32).detach().contiguous()
    runs = int((vox[1:] != vox[:-1]).sum()) + 1
    return scene, vox, xyz, feats, pts, values, runs


def roofline_at_scale(dev, peak):
    """The HBM-bound kernels at full-frame sizes, timed with CUDA events around the C-ABI calls on the REAL sample stream
    of the C3 scene (ray-marched, ray-major: ~8 consecutive samples per voxel, face-adjacent voxel transitions) and,
    for comparison, on round 1's synthetic stream (an independent random voxel every 6 samples).  Algorithmic bytes per
    unit are SURVEY.md §8d's figures."""
    from nsvf_b200 import _lib
    L, p = _lib.load(), _lib.ptr
    scene, vox, xyz, feats, pts, values, runs = real_sample_stream(dev)
    M = vox.numel()
    out = torch.empty(M, 32, device=dev)
    gv = torch.zeros_like(values)
    st = torch.cuda.current_stream().cuda_stream
    res = {"stream": "C3 scene, 800x800 view, first 64 sample columns in ray-major order: %d samples, %.1f samples per "
                     "voxel run" % (M, M / runs)}

    def rec(name, fn, nbytes):
        ms = _time(fn, n=5, warm=2)
        res[name] = {"ms": ms, "achieved_GBs": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / peak, 4),
                     "algorithmic_bytes": nbytes}
    rec("trilinear_fwd (real stream, 144 B/sample)",
        lambda: L.nsvf_trilinear_embed_fwd(st, M, 32, p(vox), p(xyz), p(feats), p(pts), p(values), scene.voxel_size, p(out)),
        M * 144)
    rec("trilinear_bwd (real stream, 144 B/sample)",
        lambda: L.nsvf_trilinear_embed_bwd(st, M, 32, p(vox), p(xyz), p(feats), p(pts), p(values), scene.voxel_size,
                                           p(out), p(gv), None), M * 144)
    g = torch.Generator(device=dev).manual_seed(0)
    vox2 = torch.randint(0, scene.n, (M // 6 + 1,), device=dev, generator=g).repeat_interleave(6)[:M].int().contiguous()
    xyz2 = (pts[vox2.long()] + (torch.rand(M, 3, device=dev, generator=g) - 0.5) * scene.voxel_size).contiguous()
    rec("trilinear_fwd (synthetic stream of round 1)",
        lambda: L.nsvf_trilinear_embed_fwd(st, M, 32, p(vox2), p(xyz2), p(feats), p(pts), p(values), scene.voxel_size, p(out)),
        M * 144)
    rec("trilinear_bwd (synthetic stream of round 1)",
        lambda: L.nsvf_trilinear_embed_bwd(st, M, 32, p(vox2), p(xyz2), p(feats), p(pts), p(values), scene.voxel_size,
                                           p(out), p(gv), None), M * 144)
    del out, xyz, vox, xyz2, vox2
    B, K = 262144, 256
    fe = torch.rand(B, K, device=dev) * 0.1
    tex = torch.rand(B, K, 3, device=dev)
    dep = torch.rand(B, K, device=dev)
    probs, od, om, oc = (torch.empty(B, K, device=dev), torch.empty(B, device=dev), torch.empty(B, device=dev),
                         torch.empty(B, 3, device=dev))
    rec("composite_fwd (dense rows, 262144 x 256, 28 B/sample)",
        lambda: L.nsvf_composite_fwd(st, B, K, p(fe), p(tex), p(dep), p(probs), p(od), p(om), p(oc)), B * K * 28 + B * 20)
    gfe, gtex = torch.empty(B, K, device=dev), torch.empty(B, K, 3, device=dev)
    rec("composite_bwd (dense rows, 262144 x 256, 36 B/sample)",
        lambda: L.nsvf_composite_bwd(st, B, K, p(fe), p(tex), p(dep), None, p(od), p(om), p(oc), p(gfe), p(gtex)),
        B * K * 36 + B * 20)
    return res


def stage_times(dev, pipe, batch):
    """Informational per-stage device times of one C2 step (separate, untimed pass)."""
    from nsvf_b200 import clib, ops
    enc = pipe.encoder
    rs, rd, _ = batch
    st = enc.precompute(id=torch.zeros(1, dtype=torch.long, device=dev))
    pts = st["voxel_center_xyz"]
    S, V, P, _ = rd.shape
    rs_f = rs.expand_as(rd).contiguous().view(S, V * P, 3)
    rd_f = rd.reshape(S, V * P, 3)
    res = {}
    res["intersect_kernel_call"] = _time(lambda: clib.aabb_ray_intersect(enc.voxel_size, enc.max_hits, pts, rs_f, rd_f))
    res["ray_intersect_total(sort+mask glue)"] = _time(lambda: enc.ray_intersect(rs, rd, st))
    pipe.train()
    with torch.no_grad():
        _, rdd, inter, hits, _ = pipe.intersecting(rs, rd, st)
    inter = {k: v.reshape(-1, *v.shape[2:])[hits.reshape(-1)] for k, v in inter.items()}
    res["inverse_cdf_sampling"] = _time(lambda: enc.ray_sample(inter))
    samples = enc.ray_sample(inter)
    sidx = samples["sampled_point_voxel_idx"]
    mask = sidx.ne(-1)
    M = int(mask.sum())
    vox = sidx[mask]
    xyz = torch.randn(M, 3, device=dev) * 0.1 + pts.reshape(-1, 3)[vox.long()]
    feats, values = st["voxel_vertex_idx"].reshape(-1, 8).int(), st["voxel_vertex_emb"].reshape(-1, 32).detach()
    res["trilinear_fwd"] = _time(lambda: ops.trilinear_embed(vox, xyz, feats, pts.reshape(-1, 3), values, enc.voxel_size))
    v2 = values.clone().requires_grad_(True)
    emb = ops.trilinear_embed(vox, xyz, feats, pts.reshape(-1, 3), v2, enc.voxel_size)
    g = torch.randn_like(emb)
    res["trilinear_bwd"] = _time(lambda: torch.autograd.grad(emb, v2, g, retain_graph=True))
    fld = pipe.field
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)

    def mlp():
        o = fld({"emb": emb.detach().requires_grad_(True), "ray": dirs})
        (o["sigma"].sum() + o["texture"].sum()).backward()
    res["field_mlp_fwd_bwd(cuBLAS)"] = _time(mlp, n=3, warm=1)
    B, K = sidx.shape
    fe = torch.rand(B, K, device=dev).requires_grad_(True)
    tex = torch.rand(B, K, 3, device=dev).requires_grad_(True)
    dep = samples["sampled_point_depth"]
    res["composite_fwd"] = _time(lambda: ops.composite(fe, tex, dep))
    o = ops.composite(fe, tex, dep)
    res["composite_bwd"] = _time(lambda: torch.autograd.grad([o[1].sum() + o[2].sum() + o[3].sum()], [fe, tex],
                                                             retain_graph=True))
    res["samples"] = M
    res["K"] = K
    return res


def live_kernel_time(fn, name, repeats=2):
    """Device time of every launch of kernel `name` inside `fn()` (library hook: one CUDA event pair per launch, recorded
    on the launching stream) -> (launches per call, total ms per call, min launch ms, max launch ms)."""
    import ctypes
    from nsvf_b200 import _lib
    L = _lib.load()
    n, tot, mn, mx = ctypes.c_int(0), ctypes.c_float(0), ctypes.c_float(0), ctypes.c_float(0)
    fn()
    torch.cuda.synchronize()
    L.nsvf_profile_begin(name.encode())
    for _ in range(repeats):
        fn()
    torch.cuda.synchronize()
    L.nsvf_profile_end(ctypes.addressof(n), ctypes.addressof(tot), ctypes.addressof(mn), ctypes.addressof(mx))
    return n.value / repeats, tot.value / repeats, mn.value, mx.value


def frame_rooflines(dev, pipe, rs, rd, peak, peak_src, frame_ms):
    """The hand-written kernels of the ray-marching path, timed LIVE inside the C3 hot-path frame (stand-in field), against
    the HBM roofline.  Algorithmic bytes per launch are SURVEY.md §8d's compulsory traffic (DESIGN.md §4 states them per
    unit); the units one launch processes come from the frame itself (rays, emitted samples, evaluated samples)."""
    from nsvf_b200.field import TrivialField
    saved, pipe.field = pipe.field, TrivialField()
    state = {}

    def frame():
        with torch.no_grad():
            state["out"] = pipe(rs, rd)
    try:
        frame()
        out = state["out"]
        rays = int(out["hits"].sum())
        ae = int(out["ae"])
        emitted = int(out["samples"]["sampled_point_count"].sum()) if "sampled_point_count" in out["samples"] else ae
        P, n_vox = int(pipe.encoder.max_hits), int(pipe.encoder.num_voxels)
        lz_idx = out["samples"].get("lazy_pts_idx") if isinstance(out["samples"], dict) else None
        valid_bins = int(lz_idx.ne(-1).sum()) if lz_idx is not None else rays * P
        all_rays = rd.numel() // 3
        table = [  # (profile name, what, bound, bytes per FRAME as a function of launches)
            ("inverse_cdf_sampling_kernel", "inverse-CDF sampler, %d rays -> %d samples" % (rays, emitted), "hbm",
             lambda n: rays * (16 * P + 4 + 4) + 12 * emitted),
            ("inverse_cdf_plan_kernel", "on-demand sampling, per-ray sample counts: the probabilities of the %d valid bins twice "
             "(two sequential passes; the second hits L1/L2), 16 bisection probes, steps and 4 trailing-loop depths in, "
             "12 B per ray out" % valid_bins, "hbm", lambda n: 4 * valid_bins + rays * (64 + 4 + 16 + 12)),
            ("inverse_cdf_stream_kernel", "on-demand sampling, resumable serial sampler: hit lists once, 12 B per evaluated "
             "sample out, 96 B of parked state per live ray and block (lower bound: blocks run ahead of the windows)", "hbm",
             lambda n: rays * 16 * P + 12 * ae + 96 * rays * n),
            ("trilinear_fwd_kernel", "trilinear interpolation fwd, %d samples in the frame's windows" % ae, "hbm",
             lambda n: 144 * ae + 0 * n),
            ("march_compact_kernel", "window compaction (12 B in + 32 B out per sample, 9 B per ray and window)", "hbm",
             lambda n: 44 * ae + 9 * rays * n),
            ("march_epilogue_kernel", "free energy + scatter + early stop (36 B per sample, 21 B per ray and window)", "hbm",
             lambda n: 36 * ae + 21 * rays * n),
            ("march_composite_fwd_kernel", "compositing over the planes (20 B per evaluated sample, 36 B per ray)", "hbm",
             lambda n: 20 * ae + 36 * rays),
            ("march_transpose_kernel", "row -> slot-major transpose (24 B per sample; lower bound: evaluated samples)", "hbm",
             lambda n: 24 * ae),
            ("aabb_intersect_sorted_kernel", "sorted aabb intersection (lattice walk), %d rays x %d voxels: rays in, hit lists out"
             % (all_rays, n_vox), "hbm", lambda n: all_rays * (24 + 12 * P + 1) + 16 * n_vox),
        ]
        res = []
        for name, what, bound, nbytes in table:
            launches, ms, mn, mx = live_kernel_time(frame, name)
            if launches == 0 or ms <= 0:
                continue
            b = nbytes(launches)
            ach = b / (ms / 1e3) / 1e9
            res.append({"kernel": name, "what": what, "bound": bound, "achieved": round(ach, 1), "peak": peak,
                        "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": None, "peak_source": peak_src,
                        "launches_per_frame": launches, "ms_per_frame": round(ms, 4),
                        "avg_launch_ms": round(ms / launches, 5), "algorithmic_bytes_per_launch": int(b / launches),
                        "share_of_frame": round(ms / frame_ms, 4), "workload": "C3 800x800 frame, hot path only"})
        return res
    finally:
        pipe.field = saved


def frame_bench(dev, pipe=None, rs=None, rd=None):
    """ms per 800x800 frame on the C3 scene (BASELINE.json configs[2]): eval, early termination 0.01, chunk 512."""
    from nsvf_b200 import synthetic
    from nsvf_b200.field import TrivialField
    res = {}
    if rs is None:
        rs, rd = synthetic.camera_rays(RES, RES, 1, radius=4.5, seed=7, device=dev)
        rs, rd = rs[None, :, None, 0, :].contiguous(), rd[None].contiguous()
    if pipe is None:
        pipe, _ = build_model(dev, "C3", train=False, field="mlp", tolerance=0.01, chunk=512, sigma_bias=2.0)
    mlp_field = pipe.field
    for name, field in (("with_field_mlp", mlp_field), ("hot_path_only(trivial field)", TrivialField())):
        pipe.field = field
        with torch.no_grad():
            for _ in range(2):          # two warm frames: cuBLAS picks its kernels for the frame's chunk shapes
                pipe(rs, rd)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 3 if field is mlp_field else 5
            for _ in range(n):
                out = pipe(rs, rd)
            e1.record()
            torch.cuda.synchronize()
        res[name] = {"ms_per_800x800_frame": round(e0.elapsed_time(e1) / n, 3), "field_evaluations": int(out["ae"]),
                     "rays_hit": int(out["hits"].sum())}
        if field is not mlp_field:
            res["_hot_out"] = {k: out[k].detach() for k in ("colors", "depths", "missed")}
    pipe.field = mlp_field
    res["voxels"] = int(pipe.encoder.num_voxels)
    return res


# ------------------------------------------------------------------------------------------------------
# reference on the SAME B200: the unmodified reference (baseline/_ref) in a child process, on our tensors
# ------------------------------------------------------------------------------------------------------
def _ours_clib(dev, pipe, rs, rd, march, n, warm):
    """encoder.ray_intersect -> probs/steps -> ray_sample on our kernels, timed like baseline/ref_gpu.py:leg_clib."""
    enc = pipe.encoder
    with torch.no_grad():
        st = enc.precompute(id=torch.zeros(1, dtype=torch.long, device=dev))
        t_int = _time(lambda: enc.ray_intersect(rs, rd, st), n=n, warm=warm)
        _, _, inter, hits = enc.ray_intersect(rs, rd, st)
        if march is None:
            sel = hits.reshape(-1)
        else:
            sel = torch.zeros_like(hits.reshape(-1))
            sel[march.to(dev)] = True
        sub = {k: v.reshape(-1, v.size(-1))[sel] for k, v in inter.items()}

        def sample():
            o = dict(sub)
            dists = (o["max_depth"] - o["min_depth"]).masked_fill(o["intersected_voxel_idx"].eq(-1), 0)
            o["probs"] = dists / dists.sum(dim=-1, keepdim=True)
            o["steps"] = dists.sum(-1) / enc.step_size
            return enc.ray_sample(o)
        was = enc.training
        enc.eval()
        t_smp = _time(sample, n=n, warm=warm)
        enc.train(was)
    rays = rd.numel() // 3
    return {"rays_intersected": rays, "rays_sampled": int(sel.sum()), "intersect_ms": t_int, "sample_ms": t_smp,
            "rays_per_s": round(rays / ((t_int + t_smp) / 1e3), 1)}


def ref_gpu_leg(dev, pipe_c2, host_batches, march, ours, steps, c3):
    """Runs baseline/ref_gpu.py (unmodified reference, child process) on this GPU with OUR voxels, weights, rays and
    targets, then our side of the same legs.  `ours` carries the numbers the main arm already measured.
    -> (ref_gpu, vs_ref_gpu)."""
    import subprocess
    import tempfile
    from nsvf_b200 import checkpoint, synthetic
    from baseline import install_ref
    if not install_ref.installed():
        return {"unavailable": "baseline/_ref is not populated (python baseline/install_ref.py where /root/reference exists)"}, None
    tmp = tempfile.mkdtemp(prefix="nsvf_refgpu_")
    pipe_c3, scene3, rs3, rd3, fr = c3["pipe"], c3["scene"], c3["rs"].cpu(), c3["rd"].cpu(), c3["frame"]
    inp = {"steps": steps,
           "C2": {"state": {k: v.cpu() for k, v in checkpoint.to_reference_state_dict(pipe_c2).items()},
                  "bbox_line": "-1.2 -1.2 -1.2 1.2 1.2 1.2 0.4", "max_hits": 60, "chunk": 64, "tolerance": 0.0,
                  "pixel_per_view": PIX_PER_VIEW, "H": RES, "W": RES, "views": VIEWS,
                  "batches": [tuple(t.cpu() for t in b) for b in host_batches],
                  "rs": host_batches[0][0].cpu(), "rd": host_batches[0][1].cpu(), "march": march.cpu()},
           "C3": {"state": {k: v.cpu() for k, v in checkpoint.to_reference_state_dict(pipe_c3).items()},
                  "bbox_line": "-2.4 -2.4 -2.4 2.4 2.4 2.4 0.4", "max_hits": scene3.max_hits, "chunk": 512,
                  "tolerance": 0.01, "rs": rs3, "rd": rd3}}
    torch.save(inp, os.path.join(tmp, "inputs.pt"))
    env = {k: v for k, v in os.environ.items() if k not in ("CUBLAS_EMULATE_SINGLE_PRECISION", "RANK", "WORLD_SIZE",
                                                             "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    env["CUDA_VISIBLE_DEVICES"] = os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[dev.index] \
        if os.environ.get("CUDA_VISIBLE_DEVICES") else str(dev.index)
    torch.cuda.empty_cache()
    out_json = os.path.join(tmp, "ref.json")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_gpu.py"), "--inputs",
                        os.path.join(tmp, "inputs.pt"), "--out", out_json], env=env, capture_output=True, text=True,
                       timeout=900)
    if r.returncode != 0 or not os.path.exists(out_json):
        return {"error": (r.stderr or r.stdout)[-600:]}, None
    ref = json.load(open(out_json))
    ref["how"] = ("unmodified reference (baseline/_ref: NSVFModel._forward / SparseVoxelEncoder / VolumeRenderer / clib "
                  "wrappers + clib kernels built by its own setup.py for sm_100a) in a child process on this GPU, same "
                  "voxels, weights, rays and targets; CUDA events after warm-up")
    # ---- our side of the same legs ----
    mine = dict(ours)
    rs2, rd2 = host_batches[0][0].to(dev), host_batches[0][1].to(dev)
    mine["clib_C2"] = _ours_clib(dev, pipe_c2, rs2, rd2, march, 5, 2)
    mine["clib_C3"] = _ours_clib(dev, pipe_c3, rs3.to(dev), rd3.to(dev), None, 5, 2)
    mine["frame"], mine["frame_hot_path"] = fr["with_field_mlp"], fr["hot_path_only(trivial field)"]
    # parity of the two arms on the tensors that were timed (deterministic eval frame, trivial field)
    try:
        theirs = torch.load(os.path.splitext(out_json)[0] + "_frame.pt")
        o = fr["_hot_out"]
        ref["frame_parity_vs_ours"] = {
            "colors_max_abs_diff": float((o["colors"].cpu() - theirs["colors"]).abs().max()),
            "depths_max_abs_diff": float((o["depths"].cpu() - theirs["depths"]).abs().max()),
            "missed_max_abs_diff": float((o["missed"].cpu() - theirs["missed"]).abs().max()),
            "rays_with_colour_diff_above_1e-5": int(((o["colors"].cpu() - theirs["colors"]).abs().max(-1)[0] > 1e-5).sum()),
            "rays": int(theirs["missed"].numel()),
            "note": "early termination compares a running fp32 free-energy sum with -log(0.01) after every window; the two "
                    "arms interpolate with different (tolerance-equal) op orders, so a ray whose sum lands within an ulp of "
                    "the threshold can stop one window apart — those rays carry the whole difference",
            "field_evaluations_equal": bool(fr["hot_path_only(trivial field)"]["field_evaluations"]
                                            == ref.get("frame_hot_path", {}).get("field_evaluations"))}
    except Exception as e:
        ref["frame_parity_vs_ours"] = {"error": repr(e)[:200]}

    def ratio(a, b):
        try:
            return round(a / b, 2)
        except Exception:
            return None
    g = lambda d, *ks: (d.get(ks[0], {}) or {}).get(ks[1]) if len(ks) == 2 else d.get(ks[0])
    vs = {"clib_C2": ratio(g(mine, "clib_C2", "rays_per_s"), g(ref, "clib_C2", "rays_per_s")),
          "clib_C3": ratio(g(mine, "clib_C3", "rays_per_s"), g(ref, "clib_C3", "rays_per_s")),
          "step": ratio(g(ref, "step", "ms_per_step"), g(mine, "step", "ms_per_step")),
          "step_hot_path": ratio(g(ref, "step_hot_path", "ms_per_step"), g(mine, "step_hot_path", "ms_per_step")),
          "frame": ratio(g(ref, "frame", "ms_per_frame"), g(mine, "frame", "ms_per_800x800_frame")),
          "frame_hot_path": ratio(g(ref, "frame_hot_path", "ms_per_frame"),
                                  g(mine, "frame_hot_path", "ms_per_800x800_frame"))}
    ref["ours"] = mine
    return ref, vs


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the box's host cores (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------------
def cpu_arm(steps, warmup, fraction):
    import oracle
    from oracle import wrappers
    from nsvf_b200 import synthetic
    from oracle.field_ref import ReferenceRadianceField      # the reference's plain-torch composition of the MLP
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    oracle.build()
    scene = synthetic.make_scene("C2")
    pts = scene.points.copy()
    pts[:, 0] += np.float32(scene.voxel_size / 10)
    torch.manual_seed(0)
    field = ReferenceRadianceField()
    values = torch.from_numpy(scene.values.copy()).requires_grad_(True)
    opt = torch.optim.Adam(list(p for p in field.parameters() if p.requires_grad) + [values], lr=1e-3)
    pix = RES * RES // fraction
    k = PIX_PER_VIEW // fraction
    rs_all, rd_all = synthetic.camera_rays(RES, RES, VIEWS, radius=3.2, seed=0)
    rng = np.random.RandomState(0)

    def one_step():
        sel = rng.choice(RES * RES, pix, replace=False)
        rd = rd_all[:, sel].numpy()
        rs = np.broadcast_to(rs_all.numpy(), (VIEWS, RES * RES, 3))[:, sel]
        idx, dmin, dmax = oracle.aabb_intersect(rs.reshape(1, -1, 3), rd.reshape(1, -1, 3), pts, scene.voxel_size, scene.max_hits)
        idx_t, dmin_t, dmax_t, hits = wrappers.sort_hits(*[torch.from_numpy(a[0]) for a in (idx, dmin, dmax)])
        hit_ids = hits.nonzero()[:, 0].numpy()
        take = np.sort(rng.choice(hit_ids, min(VIEWS * k, len(hit_ids)), replace=False))
        idx_t, dmin_t, dmax_t = idx_t[take], dmin_t[take], dmax_t[take]
        probs, stp = wrappers.probs_and_steps(idx_t, dmin_t, dmax_t, scene.step_size)
        sidx, sdep, sdist = wrappers.inverse_cdf_sampling(wrappers.NumpyExt(), idx_t, dmin_t, dmax_t, probs, stp, -1, False)
        sidx, sdep, sdist = wrappers.mask_samples(sidx, sdep, sdist)
        mask = sidx.ne(-1)
        o = torch.from_numpy(rs.reshape(-1, 3)[take].copy())
        d = torch.from_numpy(rd.reshape(-1, 3)[take].copy())
        xyz = (o[:, None] + d[:, None] * sdep[..., None])[mask]
        vox = sidx[mask].numpy()
        emb = torch.from_numpy(oracle.trilinear_fwd(vox, xyz.numpy(), scene.feats, pts, values.detach().numpy(),
                                                    scene.voxel_size)).requires_grad_(True)
        out = field({"emb": emb, "ray": d[:, None].expand(*sdep.shape, 3)[mask]})
        fe_c = torch.relu(out["sigma"] + torch.randn_like(out["sigma"])) * sdist[mask] * 7.0
        fe = torch.zeros_like(sdep).masked_scatter(mask, fe_c)
        tex = torch.zeros(*sdep.shape, 3).masked_scatter(mask[..., None].expand(-1, -1, 3), out["texture"])
        p_, dep_, mis_, col_ = [torch.from_numpy(a) for a in oracle.composite_fwd(fe.detach().numpy(), tex.detach().numpy(), sdep.numpy())]
        target = torch.rand(col_.shape) * 2 - 1
        col_.requires_grad_(True); mis_.requires_grad_(True)
        loss = ((col_ + mis_[:, None] - target) ** 2).mean() * 128 + (mis_ ** 2).mean()
        loss.backward()
        g_fe, g_tex = oracle.composite_bwd(fe.detach().numpy(), tex.detach().numpy(), sdep.numpy(), None, None,
                                           mis_.grad.numpy(), col_.grad.numpy())
        opt.zero_grad(set_to_none=True)
        torch.autograd.backward([fe, tex], [torch.from_numpy(g_fe), torch.from_numpy(g_tex)])
        gv, _ = oracle.trilinear_bwd(vox, xyz.numpy(), scene.feats, pts, values.detach().numpy(), scene.voxel_size,
                                     emb.grad.numpy())
        values.grad = torch.from_numpy(gv)
        opt.step()
        return len(take), int(mask.sum())

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    rays = 0
    for _ in range(steps):
        r, m = one_step()
        rays += r
    dt = time.perf_counter() - t0
    value = rays / dt
    sample = ("1/%d of the C2 step per CPU step: %d rays intersected, %d rays marched (%d samples), fwd+bwd+Adam; "
              "C oracle with OpenMP for intersect/sample/interp/composite, torch-CPU fp32 MLP" %
              (fraction, VIEWS * pix, r, m))
    return {"value": value, "ms_per_step": dt / steps * 1e3, "steps": steps,
            "cpu_baseline": {"value": round(value, 2), "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_arm(max(args.steps, 1), max(args.warmup, 1), args.cpu_fraction)
    line = {"impl": "reference", "metric": METRIC, "value": round(r["value"], 2), "unit": "rays/s", "n_gpus": args.gpus,
            "steps": max(args.steps, 1), "warmup": max(args.warmup, 1), "ms_per_step": round(r["ms_per_step"], 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: nsvf_base training step (bounded CPU sample, see cpu_baseline.sample)"},
            "cpu_baseline": r["cpu_baseline"],
            "e2e": {"value": round(r["value"], 2), "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# configs[2..4] on N GPUs (same launch contract; the driver's own runs use the default C2)
# ------------------------------------------------------------------------------------------------------
def _dist_setup():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    presize_allocator(dev, large_gib=48)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return dist, world, rank, local, dev


def _timed_steps(dist, world, dev, step, n):
    """n calls of step(i) between barrier + synchronize on both sides; CUDA-event time, MAX over ranks (ms)."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for i in range(n):
        last = step(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, last


def _base_line(args, world, value, unit, ms, n_warm, workload, extra_cfg, clocks, launches, e2e):
    line = {"metric": METRIC if args.config == "C5" else "rays/s (intersect+sample+composite)", "value": round(value, 1),
            "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": dict({"workload": workload}, **extra_cfg), "clocks": clocks,
            "gpu_launches": int(launches), "e2e": e2e}
    return line


def run_c3(args):
    """configs[2]: 800x800 full-frame rendering of the ~112k-voxel scene with early termination, frames sharded over the
    ranks (fairnr_cli/render_multigpu.py:95-104: every rank renders its own frames), colours gathered on rank 0."""
    from nsvf_b200 import _lib, synthetic
    from nsvf_b200.field import TrivialField
    L = _lib.load()
    dist, world, rank, local, dev = _dist_setup()
    pipe, scene = build_model(dev, "C3", train=False, field="mlp", tolerance=0.01, chunk=512, sigma_bias=2.0)
    frames = []
    for f in range(4):     # this rank's cameras (frame f of rank r = camera 4 r + f of the trajectory)
        rs, rd = synthetic.camera_rays(RES, RES, 1, radius=4.5, seed=100 + 4 * rank + f)
        frames.append((rs[None, :, None, 0, :].contiguous().pin_memory(), rd[None].contiguous().pin_memory()))
    resident = [(a.to(dev), b.to(dev)) for a, b in frames]
    staging = (torch.empty_like(resident[0][0]), torch.empty_like(resident[0][1]))
    gathered = [torch.empty(RES * RES, 3, device=dev) for _ in range(world)] if rank == 0 else None
    host_out = torch.empty(RES * RES, 3).pin_memory()

    def render(i, from_host=False):
        if from_host:
            staging[0].copy_(frames[i % 4][0], non_blocking=True)
            staging[1].copy_(frames[i % 4][1], non_blocking=True)
            rs, rd = staging
        else:
            rs, rd = resident[i % 4]
        with torch.no_grad():
            out = pipe(rs, rd)
        if world > 1:        # deliver the frame to rank 0 (NCCL gather over NVLink)
            dist.gather(out["colors"], gathered, dst=0)
        if from_host:
            host_out.copy_(out["colors"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return out
    res = {}
    sampler = ClockSampler(local)
    sampler.start()
    n_warm = max(args.warmup, 3)
    for name, field in (("with_field_mlp", pipe.field), ("hot_path_only", TrivialField())):
        pipe.field = field
        _timed_steps(dist, world, dev, render, n_warm)
        sampler.reset()
        l0 = L.nsvf_kernel_launches()
        ms, out = _timed_steps(dist, world, dev, render, args.steps)
        res[name] = {"ms": ms, "launches": L.nsvf_kernel_launches() - l0, "ae": int(out["ae"]), "clocks": None}
        if name == "with_field_mlp":
            clocks = sampler.stop()
            ms_e2e, _ = _timed_steps(dist, world, dev, lambda i: render(i, True), args.steps)
    rays = RES * RES
    m = res["with_field_mlp"]
    line = _base_line(
        args, world, world * rays * args.steps / (m["ms"] / 1e3), "rays/s", m["ms"], n_warm,
        "C3: 800x800 full-frame render, %d voxels (voxel 0.1, step 1/8, max_hits 135), eval, early termination 0.01, "
        "chunk 512, one frame per GPU and step, frames sharded over the GPUs, field MLP fp32 on cuBLAS" % scene.n,
        {"mlp_gemm": _blas.mode(), "field_evaluations_last_frame": m["ae"], "l2": "inputs and intermediates of a frame "
         "(> 10 GB) exceed the 126 MB L2", "parallelism": "frames sharded x%d, voxel set replicated, NCCL gather of the "
         "colours to rank 0" % world},
        clocks, m["launches"],
        {"value": round(world * rays * args.steps / (ms_e2e / 1e3), 1), "unit": "rays/s",
         "h2d_bytes_per_step": sum(t.numel() * 4 for t in frames[0]), "d2h_bytes_per_step": rays * 12})
    line["ms_per_800x800_frame"] = round(m["ms"] / args.steps, 3)
    h = res["hot_path_only"]
    line["value_hot_path"] = {"value": round(world * rays * args.steps / (h["ms"] / 1e3), 1), "unit": "rays/s",
                              "ms_per_800x800_frame": round(h["ms"] / args.steps, 3), "field": "trivial (no contraction)",
                              "field_evaluations_last_frame": h["ae"]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def run_c4(args):
    """configs[3]: --use-octree intersection of a frame's rays with the ~0.9 M-voxel scene (frames sharded over the ranks)
    and one pruning pass (8^3 lattice points per voxel) with the voxels sharded over the ranks and the keep mask
    all-gathered over NCCL (fairnr/modules/encoder.py:605-618 recomputes the full mask on every rank)."""
    from nsvf_b200 import _lib, synthetic
    from nsvf_b200.encoder import SparseVoxelEncoder
    from nsvf_b200.field import RadianceField
    L = _lib.load()
    dist, world, rank, local, dev = _dist_setup()
    torch.manual_seed(0)
    scene = synthetic.make_scene("C4")
    enc = SparseVoxelEncoder(scene.points, scene.voxel_size, max_hits=scene.max_hits, use_octree=True).to(dev).eval()
    field = RadianceField(sigma_bias=0.0).to(dev).eval()
    st = enc.precompute(id=torch.zeros(1, dtype=torch.long, device=dev))        # builds the octree (host, once)
    frames = []
    for f in range(2):
        rs, rd = synthetic.camera_rays(RES, RES, 1, radius=4.5, seed=200 + 2 * rank + f, device=dev)
        frames.append((rs[None, :, None, 0, :].contiguous(), rd[None].contiguous()))
    hits_n = [0]

    def intersect(i):
        rs, rd = frames[i % 2]
        with torch.no_grad():
            _, _, inter, hits = enc.ray_intersect(rs, rd, st)
        hits_n[0] = hits
        return inter
    sampler = ClockSampler(local)
    sampler.start()
    n_warm = max(args.warmup, 3)
    _timed_steps(dist, world, dev, intersect, n_warm)
    sampler.reset()
    l0 = L.nsvf_kernel_launches()
    ms, inter = _timed_steps(dist, world, dev, intersect, args.steps)
    launches = L.nsvf_kernel_launches() - l0
    clocks = sampler.stop()
    rays = RES * RES
    # pruning: voxels sharded (64-voxel chunks stay whole, so every field call sees the rows it would see on one GPU)
    field_fn = lambda inp, outputs: field(inp, outputs=outputs)
    keep0 = enc.keep.clone()
    ms_prune, _ = _timed_steps(dist, world, dev, lambda i: (enc.keep.copy_(keep0),
                                                          enc.pruning(field_fn, th=0.5, voxel_shard=(rank, world), bits=8)), 1)
    sharded_keep = enc.keep.clone()
    prune = {"voxels": scene.n, "lattice_points_per_voxel": 512, "ms_per_pass": round(ms_prune, 2),
             "kept": int(sharded_keep.sum()), "keep_mask_allgather_bytes": scene.n,
             "collective": "NCCL all_gather of the uint8 keep mask" if world > 1 else "none (1 GPU)"}
    if world > 1 and not args.no_stages:
        # parity of the exchange: the gathered mask must equal the mask one GPU computes alone (rank 0 recomputes it)
        if rank == 0:
            enc.keep.copy_(keep0)
            enc.pruning(field_fn, th=0.5, bits=8)
            prune["mask_equals_single_rank"] = bool(torch.equal(enc.keep, sharded_keep))
        dist.barrier()
    line = _base_line(
        args, world, world * rays * args.steps / (ms / 1e3), "rays/s", ms, n_warm,
        "C4: --use-octree svo_ray_intersect + sort of an 800x800 frame's rays per GPU against %d voxels (%d octree nodes, "
        "max_hits 202), then one pruning pass (8^3 points per voxel) with the voxels sharded and the keep mask "
        "all-gathered" % (scene.n, int(enc.flatten_centers.shape[0])),
        {"l2": "hit lists of a frame (1.5 GB) exceed the 126 MB L2", "rays_hit_last_frame": int(hits_n[0].sum()),
         "parallelism": "frames sharded x%d for the intersection; voxels sharded x%d for pruning" % (world, world)},
        clocks, launches, {"value": None, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                           "note": "kernel-level config: rays are generated on the device, no end-to-end leg"})
    line["metric"] = "rays/s (octree intersection + sort), pruning ms per pass"
    line["prune"] = prune
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def run_c5(args):
    """configs[4]: Tanks&Temples-shaped 1920x1080 scene (elongated bbox), hierarchical path (coarse pass + inverse-CDF
    importance-sampled fine pass, fairnr/models/nerf.py:64-79), training step with the NCCL gradient all-reduce."""
    from nsvf_b200 import _lib, synthetic
    from nsvf_b200.encoder import SparseVoxelEncoder
    from nsvf_b200.field import RadianceField
    from nsvf_b200.pipeline import NSVFPipeline
    from nsvf_b200.renderer import VolumeRenderer
    L = _lib.load()
    dist, world, rank, local, dev = _dist_setup()
    H, W, V = 1080, 1920, 2
    torch.manual_seed(0)
    scene = synthetic.make_scene("C5")
    enc = SparseVoxelEncoder(scene.points, scene.voxel_size, max_hits=scene.max_hits)
    pipe = NSVFPipeline(enc, RadianceField(), VolumeRenderer(chunk_size=64, discrete_regularization=True),
                        pixel_per_view=PIX_PER_VIEW, hierarchical_sampling=True, fixed_fine_num_samples=64).to(dev).train()
    params = [p for p in pipe.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3, betas=(0.9, 0.999))
    flat_grad = torch.zeros(sum(p.numel() for p in params), device=dev)
    off = 0
    for p in params:
        p.grad = flat_grad[off: off + p.numel()].view_as(p)
        off += p.numel()
    host = []
    for b in range(2):
        rs, rd = synthetic.camera_rays(H, W, V, radius=9.0, fov_focal=1111.0 * 800.0 / W * 1.6, seed=137 * b + rank)
        g = torch.Generator().manual_seed(1000 + 137 * b + rank)
        target = torch.rand(V * H * W, 3, generator=g) * 2 - 1
        host.append((rs[None, :, None, 0, :].contiguous().pin_memory(), rd[None].contiguous().pin_memory(),
                     target.pin_memory()))
    resident = [tuple(t.to(dev) for t in b) for b in host]
    staging = tuple(torch.empty_like(t, device=dev) for t in host[0])

    def step(i, from_host=False):
        if from_host:
            for d, s_ in zip(staging, host[i % 2]):
                d.copy_(s_, non_blocking=True)
            rs, rd, target = staging
        else:
            rs, rd, target = resident[i % 2]
        out = pipe(rs, rd)
        loss = loss_fn(out, target)
        flat_grad.zero_()
        loss.backward()
        if world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG)
        opt.step()
        if from_host:
            loss.item()
        return out
    sampler = ClockSampler(local)
    sampler.start()
    n_warm = max(args.warmup, 3)
    _timed_steps(dist, world, dev, step, n_warm)
    sampler.reset()
    l0 = L.nsvf_kernel_launches()
    ms, out = _timed_steps(dist, world, dev, step, args.steps)
    launches = L.nsvf_kernel_launches() - l0
    clocks = sampler.stop()
    ms_e2e, _ = _timed_steps(dist, world, dev, lambda i: step(i, True), args.steps)
    marched = V * PIX_PER_VIEW
    line = _base_line(
        args, world, world * marched * args.steps / (ms / 1e3), "rays/s", ms, n_warm,
        "C5: Tanks&Temples-shaped training step, %d voxels (bbox 12 x 8 x 6.4, voxel 0.2, step 1/8), %d views x 1920x1080 rays "
        "intersected + %d x 2048 rays marched per GPU, hierarchical: coarse pass + 64 importance samples per ray "
        "(inverse_cdf_sampling on the coarse samples), fwd+bwd+Adam, field MLP fp32 on cuBLAS" % (scene.n, V, V),
        {"mlp_gemm": _blas.mode(), "samples_evaluated_per_step": int(out["ae"]), "l2": "per-step intermediates exceed the "
         "126 MB L2 (2 x 2 M rays intersected)", "parallelism": "dp%d (rays sharded by view, voxel set replicated, one "
         "flat-bucket NCCL grad all-reduce per step)" % world},
        clocks, launches,
        {"value": round(world * marched * args.steps / (ms_e2e / 1e3), 1), "unit": "rays/s",
         "h2d_bytes_per_step": sum(t.numel() * 4 for t in host[0]), "d2h_bytes_per_step": 4})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        {"C2": run_ours, "C3": run_c3, "C4": run_c4, "C5": run_c5}[a.config](a)
