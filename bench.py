#!/usr/bin/env python
"""bench.py -- headline benchmark of the render hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2|stock|cfg3|cfg5]

Metric (BASELINE.json): Mrays/s, path trace 420x420, -m 1 -N 128.  A "step" is one full pass of the hot
path over the frame: every pixel sample of the 420x420, N=128 path trace on the example.off Cornell
scene (BASELINE configs[1], 11 666 triangles) -- 22.6 M samples, ~2.3e8 logical rays.  A "ray" is one
logical RayTracer::rayTrace invocation of the reference algorithm (primary, bounce, shadow), counted
by device counters.

N > 1 (launched by torchrun, one rank per GPU): weak scaling by sample index -- every rank renders
its own 128 samples per pixel of an N*128-sample frame (the stratum depends on the global sample
index, source/RayTracer.h:111-115), then the fp32 framebuffer sum and the hit counters are reduced
to rank 0 with NCCL inside the timed region.

--impl reference: the reference's own CPU implementation (the unmodified sources compiled into
oracle/_ref, else the CPU restatement) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD_SCENES = os.path.join(ROOT, "tests", "golden", "scenes")

WORKLOADS = {
    # name: (scene file, width, height, N per rank, mode, photons, k, description)
    "cfg2": ("example", 420, 420, 128, 1, 0, 0,
             "path trace -m 1 -N 128 420x420, example.off in the Cornell box (11666 tris), 3 area lights, microfacet"),
    "stock": ("stock", 420, 420, 128, 1, 0, 0, "path trace -m 1 -N 128 420x420, stock two-cube Cornell scene (34 tris)"),
    "cfg3": ("example", 420, 420, 128, 1, 50000, 10, "cfg2 + photon map -p 50000 -k 10"),
    "cfg1": ("lowres", 420, 420, 1, 0, 0, 0, "ray trace -m 0 -N 1 420x420, example_low_res.off (1222 tris)"),
}


def bytes_per_ray(num_triangles: int) -> int:
    """48 (ray in + hit out) + 64 * ceil(log2 T) (both child boxes per level) + 48 (one triangle)."""
    return 96 + 64 * math.ceil(math.log2(max(num_triangles, 2)))


def bytes_per_query(photons: int, k: int) -> int:
    return 28 + 16 * (math.ceil(math.log2(max(photons, 2))) + k) + 12 * k


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------- CPU arm
def _ref_worker(args):
    """One process of the reference arm: rows [y0,y1) of the window through oracle/_ref (unmodified reference)."""
    scene_file, w, h, N, mode, seed, x0, y0, x1, y1, s0, s1, kind = args
    from oracle import oracle as O
    flat = O.FlatScene.load(scene_file)
    flat.w, flat.h = w, h
    if kind == "reference":
        o = O.RefOracle()
        o.set_scene(flat)
    else:
        o = O.PortOracle(flat)
    t = time.perf_counter()
    o.render(N, mode, seed, window=(x0, y0, x1, y1), samples=(s0, s1))
    return time.perf_counter() - t


def cpu_sample(workload, cores, seconds_target):
    """Choose a centred window x 2 samples of the workload worth ~seconds_target on `cores` cores."""
    scene_name, w, h, N, mode, photons, k, _ = WORKLOADS[workload]
    import numpy as np
    from oracle import oracle as O
    flat = O.FlatScene.load(os.path.join(GOLD_SCENES, f"{scene_name}.rtscene"))
    per_ray = 24e-9 * flat.T + 0.3e-6  # ~24 ns per triangle per ray, brute force
    samples = max(64, int(cores * seconds_target / (10.5 * per_ray)))
    side = int(min(h, w, max(8, math.sqrt(samples / 2))))
    side -= side % 2
    x0, y0 = (w - side) // 2, (h - side) // 2
    return dict(scene_file=os.path.join(GOLD_SCENES, f"{scene_name}.rtscene"), w=w, h=h, N=N, mode=mode,
                window=(x0, y0, x0 + side, y0 + side), samples=(0, 2), T=flat.T)


def run_cpu(workload, cores, seconds_target, repeats=1, warm=0):
    """Times the reference CPU path on all `cores` (one process per core, rows interleaved by band).
    Returns (Mrays/s, description, kind, rays, seconds list)."""
    import multiprocessing as mp
    from oracle import oracle as O
    kind = "reference" if O.have_ref() else "port"
    cs = cpu_sample(workload, cores, seconds_target)
    x0, y0, x1, y1 = cs["window"]
    rows = y1 - y0
    nproc = max(1, min(cores, rows))
    bands = [(y0 + rows * i // nproc, y0 + rows * (i + 1) // nproc) for i in range(nproc)]
    jobs = [(cs["scene_file"], cs["w"], cs["h"], cs["N"], cs["mode"], 1, x0, a, x1, b, cs["samples"][0],
             cs["samples"][1], kind) for a, b in bands if b > a]
    # logical ray count of exactly this sample, from the restatement's counters (identical streams,
    # bit-identical control flow -- tests/test_oracle.py), outside the timed region
    flat = O.FlatScene.load(cs["scene_file"])
    flat.w, flat.h = cs["w"], cs["h"]
    port = O.PortOracle(flat)
    port.counters(reset=True)
    port.render(cs["N"], cs["mode"], 1, window=cs["window"], samples=cs["samples"], threads=cores)
    rays = port.counters(reset=True)["rays"]
    times = []
    ctx = mp.get_context("fork")
    with ctx.Pool(len(jobs)) as pool:
        for it in range(warm + repeats):
            t = time.perf_counter()
            pool.map(_ref_worker, jobs)
            dt = time.perf_counter() - t
            if it >= warm:
                times.append(dt)
    desc = (f"{cs['window'][2] - cs['window'][0]}x{cs['window'][3] - cs['window'][1]} px centred window x samples "
            f"[0,2) of the {cs['w']}x{cs['h']} N={cs['N']} frame = {rays} rays, {len(jobs)} processes")
    return rays, times, desc, kind, len(jobs)


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    _, w, h, N, mode, photons, k, wdesc = WORKLOADS[a.workload]
    rays, times, desc, kind, nproc = run_cpu(a.workload, cores, 6.0, repeats=a.steps, warm=a.warmup)
    total = sum(times)
    value = rays * len(times) / total / 1e6
    line = {"impl": "reference", "metric": "Mrays/s (path trace 420x420, N=128)", "value": value, "unit": "Mrays/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wdesc, "width": w, "height": h, "N": N, "mode": mode, "step": desc},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": nproc, "kind": kind, "sample": desc},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    a = ap.parse_args()
    if a.impl == "reference":
        return reference_arm(a)

    import numpy as np
    import torch
    import torch.distributed as dist
    import ray_tracing_engine_b200 as rt

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    scene_name, W, H, N_rank, mode, photons, k, wdesc = WORKLOADS[a.workload]
    scene = rt.Scene.load(os.path.join(GOLD_SCENES, f"{scene_name}.rtscene"))
    scene.w, scene.h = W, H
    N_total = N_rank * world
    kw = dict(seed=1, device=local, sample_first=rank * N_rank, sample_count=N_rank)
    r = rt.Renderer(scene, N_total, mode, None, photons, k or 5, **kw)
    if photons:
        r.build_photon_map()  # emission + kd-tree once, outside the per-step render (Renderer.cpp:209-213)
    sum_t = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    cnt_t = torch.zeros((H, W), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    background = rt.Image(W, H).fillBackground().pixels

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        """One pass of the hot path, inputs resident in HBM; returns seconds (host clock around device syncs)."""
        flush.fill_(rank)  # L2 flush between iterations (outside the timed region)
        barrier()
        t0 = time.perf_counter()
        r.render_accumulate_device(sum_t.data_ptr(), cnt_t.data_ptr())
        if world > 1:
            dist.reduce(sum_t, 0)
            dist.reduce(cnt_t, 0)
        barrier()
        return time.perf_counter() - t0

    def step_e2e():
        """The user-facing call with HOST buffers: scene upload + BVH build (rt_create), render, reduce,
        device->host read of the frame, composite on the host."""
        flush.fill_(rank)
        barrier()
        t0 = time.perf_counter()
        r2 = rt.Renderer(scene, N_total, mode, None, photons, k or 5, **kw)
        if world == 1:
            img = rt.Image(W, H)
            img.pixels = background.copy()
            r2.render(img)
        else:
            r2.render_accumulate_device(sum_t.data_ptr(), cnt_t.data_ptr())
            dist.reduce(sum_t, 0)
            dist.reduce(cnt_t, 0)
            if rank == 0:
                rt.Renderer.composite(N_total, sum_t.cpu().numpy(), cnt_t.cpu().numpy(), background)
        barrier()
        dt = time.perf_counter() - t0
        st = r2.stats()
        r2.close()
        return dt, st

    for _ in range(max(a.warmup, 0)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r.reset_stats()
    barrier()
    times, trace_ms, dev_ms = [], 0.0, 0.0
    for _ in range(a.steps):
        times.append(step())
        st = r.stats()
        trace_ms += st["trace_ms"]
        dev_ms += st["device_ms"]
    st = r.stats()
    clocks = sampler.stop() if rank == 0 else None
    # e2e (same number of steps)
    e2e_times, e2e_rays = [], 0
    for i in range(a.warmup + a.steps):
        dt, st2 = step_e2e()
        if i >= a.warmup:
            e2e_times.append(dt)
            e2e_rays += st2["rays"]

    tt = torch.tensor([sum(times), sum(e2e_times), float(st["rays"]), float(e2e_rays), trace_ms, dev_ms,
                       float(st["kernel_launches"]), float(st["knn_queries"])], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = tt
    if rank == 0:
        total_s, e2e_s = float(mx[0]), float(mx[1])
        rays_all, e2e_rays_all = float(sm[2]), float(sm[3])
        value = rays_all / total_s / 1e6
        peak, peak_src = hbm_peak()
        bpr = bytes_per_ray(scene.T)
        nq = float(sm[7])
        bq = bytes_per_query(int(st["photons_stored"]), k) if photons else 0
        launches_seg = (3 if mode == 1 else 1) * a.steps
        # dominant kernel: k_segment (3 launches per step in path mode); rank 0's launches
        alg_bytes = (float(st["rays"]) * bpr + float(st["knn_queries"]) * bq)
        seg_s = trace_ms / 1e3 if world == 1 else float(tt[4]) / 1e3
        achieved = alg_bytes / seg_s / 1e9 if seg_s > 0 else 0.0
        scene_bytes = st["bvh_nodes"] * 64 + scene.T * (48 + 16) + scene.V * 32 + scene.M * 32
        line = {
            "metric": "Mrays/s (path trace 420x420, N=128)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total_s / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wdesc, "width": W, "height": H, "N_per_gpu": N_rank, "N_total": N_total,
                       "mode": mode, "photons": photons, "k": k, "triangles": scene.T,
                       "parallelism": f"sample-index sharded x{world}, fp32 framebuffer reduced to rank 0 (NCCL)",
                       "l2": "256 MiB flush write between timed steps", "rays_per_step": rays_all / a.steps,
                       "seed": 1},
            "clocks": clocks,
            "e2e": {"value": e2e_rays_all / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes),
                    "d2h_bytes_per_step": int(W * H * 16), "ms_per_step": 1e3 * e2e_s / a.steps,
                    "includes": "rt_create (host BVH build + scene H2D), render, D2H of sums+counters, host composite"},
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "k_segment<1,false> (3 launches/step)",
                         "bytes_per_ray": bpr, "bytes_per_query": bq,
                         "alg_bytes_per_launch": alg_bytes / launches_seg,
                         "avg_launch_ms": 1e3 * seg_s / launches_seg,
                         "kernel_share_of_step": seg_s / (float(tt[0]) if float(tt[0]) > 0 else 1.0),
                         "note": "algorithmic bytes per SURVEY.md 8(d); the scene is L2-resident, the kernel is "
                                 "issue/latency bound, so a small HBM fraction is expected (see DESIGN.md)"},
        }
        if world == 1 and not a.no_cpu_baseline:
            try:
                cores = host_cores()
                rays, tms, desc, kind, nproc = run_cpu(a.workload, cores, a.cpu_seconds)
                line["cpu_baseline"] = {"value": rays / tms[0] / 1e6, "unit": "Mrays/s", "cores": nproc, "kind": kind,
                                        "sample": desc, "seconds": tms[0]}
            except Exception as e:  # the CPU leg must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port",
                                        "sample": f"failed: {e!r}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
