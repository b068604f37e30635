import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ray_tracing_engine_b200 as rt
from oracle import oracle as O
scene = rt.Scene.load("tests/golden/scenes/stock.rtscene")
flat = O.FlatScene.load("tests/golden/scenes/stock.rtscene")
port = O.PortOracle(flat)
r = rt.Renderer(scene, 128, 1, seed=1)
rb = rt.Renderer(scene, 128, 1, seed=1, flags=1)
nodes, depth = r.bvh()
print("bvh nodes", len(nodes), "depth", depth)
np.save("gpurun_out/bvh_stock_nodes.npy", nodes)
out = []
for (x, y) in ((394, 115), (256, 196), (252, 306)):
    a, fa = r.render_samples(window=(x, y, x + 1, y + 1))
    b, fb = rb.render_samples(window=(x, y, x + 1, y + 1))
    bad = np.nonzero(fa[:, 0, 0] != fb[:, 0, 0])[0]
    print("pixel", x, y, "bad samples", bad)
    for i in bad:
        jit = port.jitter(1, 1, int(i) * 420 * 420 + y * 420 + x, 1, int(i), 128)
        ray = port.camera_rays(np.array([[x, y]], np.int32), jit)
        h1, h2 = r.rayTrace(ray), r.rayTrace(ray, brute_force=True)
        hp = port.trace(ray)
        print("  ray", ray.tolist(), "\n  bvh", h1["tri_index"], h1["uvd"], "\n  brute", h2["tri_index"], h2["uvd"], "\n  port", hp["tri_index"], hp["uvd"])
        out.append(ray[0])
np.save("gpurun_out/bad_rays.npy", np.array(out))
