import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ray_tracing_engine_b200 as rt
scene = rt.Scene.load(os.path.join(ROOT, "tests/golden/scenes/stock.rtscene"))
r = rt.Renderer(scene, 16, 1, None, 50000, 10, seed=1)
r.build_photon_map()
s, c = r.render_accumulate()
print(r.stats()["device_ms"])
