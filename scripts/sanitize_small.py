"""Small invocations of every kernel family, for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py`:
direct lighting (1, 3, 5 lights), path mode, photon map (k = 10 array, k = 50 heap, k = 80 global scratch, exact mode with
the device kd build, persistent gather), device photon shards + splice, tile shards, the device BVH build, packed frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import ray_tracing_engine_b200 as rt
from ray_tracing_engine_b200 import distributed as D

def scene(name, w=48, h=36):
    s = rt.Scene.load(os.path.join(ROOT, "tests/golden/scenes", name + ".rtscene"))
    s.w, s.h = w, h
    return s

for name in ("stock", "stock_1light", "stock_5lights", "lowres"):
    for mode in (0, 1):
        r = rt.Renderer(scene(name), 3, mode, seed=2, samples_per_batch=2)
        s, c = r.render_accumulate()
        print(name, mode, float(s.sum()), int(c.sum()))
        r.close()
os.environ["RT_BVH_BUILD"] = "gpu"
r = rt.Renderer(scene("lowres"), 2, 1, seed=2)
print("device bvh", float(r.render_accumulate()[0].sum()))
r.close()
os.environ.pop("RT_BVH_BUILD")
for k, flags, env in ((10, 0, {}), (50, 0, {}), (80, 0, {}), (10, rt.RT_FLAG_KNN_EXACT, {}), (10, 0, {"RT_KNN_GATHER": "1"}), (80, 0, {"RT_KNN_GATHER": "1"})):
    os.environ.update(env)
    r = rt.Renderer(scene("stock"), 2, 1, None, 3000, k, seed=2, flags=flags, shard_rank=1, shard_count=3)
    s, c = r.render_accumulate()
    q = np.random.default_rng(1).uniform(-1, 1, (300, 3)).astype(np.float32)
    idx = r.knearest(q, k)
    print("photons k", k, flags, env, float(s.sum()), int(idx.sum()), r.stats()["photons_stored"])
    r.close()
    for e in env:
        os.environ.pop(e)
r = rt.Renderer(scene("stock"), 1, 0, None, 3000, 5, seed=5)
world, per, L = 3, r.photons_per_light(), 3
cap = max(D.path_range(per, q, world)[1] for q in range(world)) * L
gathered = torch.empty((world * cap, 7), dtype=torch.float32, device="cuda")
counts = np.zeros((world, L), np.int64)
for q in range(world):
    first, count = D.path_range(per, q, world)
    counts[q], _ = r.emit_photons_device(first, count, gathered[q * cap:].data_ptr(), cap)
total = int(counts.sum())
out = torch.empty((total, 7), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
print("splice", r.splice_photons_device(gathered.data_ptr(), world, cap, counts, out.data_ptr(), total))
r.set_photons_device(out.data_ptr(), total)
packed = torch.empty((36, 48, 4), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
r.render_accumulate_packed_device(packed.data_ptr())
print("packed", float(packed.sum()))
img = r.render(rt.Image(48, 36).fillBackground())
print("render", float(img.pixels.sum()))
rays = np.random.default_rng(2).normal(size=(500, 6)).astype(np.float32)
print("trace", int(r.rayTrace(rays)["hit"].sum()), int(r.occluded(rays).sum()), int(r.rayTrace(rays, brute_force=True)["hit"].sum()))
r.close()
print("SANITIZE_SMALL DONE")
