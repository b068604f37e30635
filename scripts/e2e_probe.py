import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, ray_tracing_engine_b200 as rt
scene = rt.Scene.load(os.path.join(ROOT, "tests/golden/scenes/example.rtscene"))
bg = rt.Image(420, 420).fillBackground().pixels
for it in range(4):
    t0 = time.perf_counter(); r = rt.Renderer(scene, 128, 1, seed=1); t1 = time.perf_counter()
    img = rt.Image(420, 420); img.pixels = bg.copy(); r.render(img); t2 = time.perf_counter()
    st = r.stats(); r.close(); t3 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms (lib create {st['create_ms']:.1f}, bvh {st['bvh_build_ms']:.1f}) render call {1e3*(t2-t1):.1f} ms (device {st['device_ms']:.1f}) close {1e3*(t3-t2):.1f} ms")
