"""Two cfg2 frames (420x420, -m 1 -N 128, example.off scene) for ncu: capture the second one
(--launch-skip = launches of the first).  Usage: profile_frame.py [cfg2|cfg3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ray_tracing_engine_b200 as rt
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
scene = rt.Scene.load(os.path.join(ROOT, "tests/golden/scenes/example.rtscene"))
scene.w = scene.h = 420
if which == "cfg3":
    r = rt.Renderer(scene, 128, 1, None, 50000, 10, seed=1)
    r.build_photon_map()
else:
    r = rt.Renderer(scene, 128, 1, seed=1)
r.reset_stats()
for i in range(2):
    r.render_accumulate()
    st = r.stats()
    print(f"frame {i}: device {st['device_ms']:.2f} ms, trace {st['trace_ms']:.2f} ms, launches so far {st['kernel_launches']}, rays {st['rays']}")
