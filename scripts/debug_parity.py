import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ray_tracing_engine_b200 as rt
g = np.load("tests/golden/render_stock_m1_N128.npz")
scene = rt.Scene.load("tests/golden/scenes/stock.rtscene")
for flags, spb in ((0, 0), (1, 0), (0, 16), (0, 1)):
    r = rt.Renderer(scene, 128, 1, seed=1, flags=flags, samples_per_batch=spb)
    s, c = r.render_accumulate()
    bad = c != g["counter"]
    print("flags", flags, "spb", spb, "counter mismatches:", int(bad.sum()), "max diff", int(np.abs(c - g["counter"]).max()),
          "rmse", float(np.sqrt(np.mean((s/128 - g["sum_rgb"]/128)**2))), "rays", r.stats()["rays"])
    if bad.any():
        ys, xs = np.nonzero(bad)
        print("  first bad pixels:", list(zip(xs[:8].tolist(), ys[:8].tolist())), c[bad][:8], g["counter"][bad][:8])
