import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, ray_tracing_engine_b200 as rt
scene = rt.Scene.load(os.path.join(ROOT, "tests/golden/scenes/stock.rtscene"))
for (N, mode, p, k) in ((128, 1, 50000, 10), (1, 0, 500000, 50)):
    r = rt.Renderer(scene, N, mode, None, p, k, seed=1)
    t = time.perf_counter(); r.build_photon_map(); tb = time.perf_counter() - t
    st0 = r.stats()
    s, c = r.render_accumulate(); r.reset_stats()
    s, c = r.render_accumulate(); st = r.stats()
    print(f"-m {mode} -N {N} -p {p} -k {k}: photon map {tb*1e3:.0f} ms (emit kernel {st0['photon_ms']:.1f} ms, kd build {st0['kd_build_ms']:.0f} ms, stored {st0['photons_stored']}); "
          f"render device {st['device_ms']:.1f} ms, rays {st['rays']}, queries {st['knn_queries']} ({st['kd_visits']/max(st['knn_queries'],1):.1f} node visits each), trace {st['trace_ms']:.1f} ms, checksum {float(s.sum()):.3f}", flush=True)
