import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, ray_tracing_engine_b200 as rt
scene = rt.Scene.load("/root/repo/tests/golden/scenes/stock.rtscene")
W = H = 96
for (N, mode, photons, k) in ((2, 0, 0, 5), (2, 0, 30000, 10), (2, 1, 30000, 10), (4, 1, 0, 5)):
    full = rt.Renderer(scene, N, mode, None, photons, k, seed=5, width=W, height=H)
    if photons: full.build_photon_map()
    plist = full.kdtree()[0] if photons else None
    fs, fc = full.render_accumulate()
    acc_s, acc_c = np.zeros_like(fs), np.zeros_like(fc)
    for first in range(N):
        r = rt.Renderer(scene, N, mode, None, photons, k, seed=5, width=W, height=H, sample_first=first, sample_count=1)
        if photons: r.build_photon_map()
        s, c = r.render_accumulate()
        acc_s += s; acc_c += c
        print("   part", first, "rays", r.stats()["rays"], "sum", float(s.sum()), "cnt", int(c.sum()))
    print(N, mode, photons, "max diff", float(np.abs(acc_s - fs).max()), "cnt equal", bool((acc_c == fc).all()), "full sum", float(fs.sum()), "rays", full.stats()["rays"])
