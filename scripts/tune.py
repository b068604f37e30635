"""Time kernel tuning variants on the headline workload (device time from CUDA events inside the library).
usage: python scripts/tune.py [variant ...]  (variants 0, 1, 5, 10)   (each variant runs in its own process: RT_TRACE_VARIANT)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np
    import ray_tracing_engine_b200 as rt
    scene_name = os.environ.get("RT_SCENE", "example")
    N = int(os.environ.get("RT_N", "128"))
    scene = rt.Scene.load(os.path.join(ROOT, "tests/golden/scenes", scene_name + ".rtscene"))
    r = rt.Renderer(scene, N, 1, seed=1)
    s, c = r.render_accumulate()
    r.reset_stats()
    dev, tr = [], []
    for _ in range(3):
        s, c = r.render_accumulate()
        st = r.stats()
        dev.append(st["device_ms"]); tr.append(st["trace_ms"])
    print(json.dumps(dict(variant=os.environ.get("RT_TRACE_VARIANT", "default"), sort=os.environ.get("RT_SORT_HITS", "default"), device_ms=min(dev), trace_ms=min(tr),
                          rays=st["rays"] // 3, grays=st["rays"] / 3 / min(dev) / 1e6, checksum=float(s.sum()), hits=int(c.sum()))))
else:
    for v in sys.argv[1:] or ["0", "1", "2", "3", "4"]:
        env = dict(os.environ)
        for kv in v.split(","):  # "1" or "1,RT_SORT_HITS=0"
            if "=" in kv:
                k_, v_ = kv.split("=")
                env[k_] = v_
            else:
                env["RT_TRACE_VARIANT"] = kv
        out = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True)
        print(out.stdout.strip() or out.stderr[-500:], flush=True)
