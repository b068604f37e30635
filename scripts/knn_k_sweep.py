"""k-NN flavour sweep: render time of a photon-gather frame for several k with the array (RT_KNN_HEAP_FROM_K=1000) and the
heap (RT_KNN_HEAP_FROM_K=0) restatement of kdtree::knearest.  Run each setting in its own process."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import ray_tracing_engine_b200 as rt
    scene = rt.Scene.load(os.path.join(ROOT, "tests/golden/scenes/stock.rtscene"))
    out = {}
    for k in (4, 10, 16, 24, 32, 50, 64):
        r = rt.Renderer(scene, 8, 1, None, 100000, k, seed=1); r.build_photon_map()
        r.render_accumulate(); r.reset_stats(); s, c = r.render_accumulate()
        out[k] = round(r.stats()["device_ms"], 2); r.close()
    print(json.dumps(out))
else:
    for name, v in (("array", "1000"), ("heap", "0")):
        o = subprocess.run([sys.executable, __file__, "--one"], env=dict(os.environ, RT_KNN_HEAP_FROM_K=v), capture_output=True, text=True)
        print(name, o.stdout.strip() or o.stderr[-300:], flush=True)
