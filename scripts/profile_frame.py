"""One frame of a bench workload bracketed by cudaProfilerStart/Stop, for ncu --profile-from-start off.

usage: profile_frame.py cfg2|cfg3|cfg4|cfg5|stock [out.json]      (RT_PROFILE_EMIT=1: the photon emission is captured too)
Prints (and writes to out.json) the frame's logical work: rays by kind, queries, launches by kernel class, and the
CUDA-event time of every class measured on an untouched warm-up frame before the capture.  scripts/ncu_roofline.py
joins this with the ncu CSV into profiles/ncu_<workload>.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
import ray_tracing_engine_b200 as rt

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
scene, W, H, N, mode, photons, k, _ = bench.load_workload(rt, which, 0)
if which == "cfg5":
    N = int(os.environ.get("RT_PROFILE_N", "16"))  # rays per sample do not depend on N; 1024 samples would be 64x longer
r = rt.Renderer(scene, N, mode, None, photons, k or 5, seed=1)
if photons:
    r.build_photon_map()
emit = r.stats()["kernel_ms"]["emit"]
r.render_accumulate()           # warm-up frame, untouched by the profiler
r.reset_stats()
r.render_accumulate()
warm = r.stats()
r.reset_stats()
r.lib.rt_profiler_range(1)
if photons and os.environ.get("RT_PROFILE_EMIT"):
    r.build_photon_map()
r.render_accumulate()
r.lib.rt_profiler_range(0)
st = r.stats()
out = dict(workload=which, width=W, height=H, N=N, mode=mode, photons=photons, k=k, triangles=scene.T,
           rays=st["rays"], primary_rays=st["primary_rays"], bounce_rays=st["bounce_rays"], shadow_rays=st["shadow_rays"],
           knn_queries=st["knn_queries"], kd_visits=st["kd_visits"], samples=st["samples"], photons_stored=st["photons_stored"],
           kernel_count=st["kernel_count"], kernel_ms_unprofiled=warm["kernel_ms"], device_ms_unprofiled=warm["device_ms"],
           emit_ms=emit)
print(json.dumps(out))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
