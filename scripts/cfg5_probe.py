"""cfg 5 probe: 1.23 M-triangle scene (example_low_res.off subdivided 5x in the Cornell box), 1920x1080."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ray_tracing_engine_b200 as rt
meshes = os.path.join(ROOT, "oracle", "_ref", "meshes")
t = time.time(); scene = rt.Scene.build(1920, 1080, meshes, os.path.join(meshes, "example_low_res.off"), 5); t_scene = time.time() - t
N = int(os.environ.get("RT_N", "16"))
t = time.time(); r = rt.Renderer(scene, N, 1, seed=1); t_create = time.time() - t
nodes, depth = r.bvh()
st = r.stats()
print(f"scene V={scene.V} T={scene.T} build {t_scene:.2f}s; rt_create {t_create:.2f}s (library {st['create_ms']:.0f} ms, of which BVH build "
      f"{st['bvh_build_ms']:.0f} ms, RT_BVH_BUILD={os.environ.get('RT_BVH_BUILD', 'default')}); bvh nodes {len(nodes)} depth {depth}", flush=True)
t = time.time(); r2 = rt.Renderer(scene, N, 1, seed=1); st2 = r2.stats(); r2.close()
print(f"second rt_create {time.time() - t:.3f}s (library {st2['create_ms']:.0f} ms, BVH build {st2['bvh_build_ms']:.0f} ms)", flush=True)
# parity at scale: BVH vs brute force on primary + secondary rays
g = np.random.default_rng(3)
n = 40000
o = np.tile(np.float32([0.3, 0.6, 2.3]), (n, 1)); d = (g.normal(size=(n, 3)) * [0.25, 0.25, 0.1] + [-0.12, -0.25, -1]).astype(np.float32)
rays = np.concatenate([o, d], 1).astype(np.float32)
a = r.rayTrace(rays); ok = a["hit"] == 1
p = (o + d * a["uvd"][:, 2:3])[ok]
rays2 = np.concatenate([p, g.normal(size=p.shape).astype(np.float32)], 1)
for name, batch in (("primary", rays), ("secondary", rays2)):
    t = time.time(); x = r.rayTrace(batch); t1 = time.time() - t
    t = time.time(); y = r.rayTrace(batch, brute_force=True); t2 = time.time() - t
    same = (x["tri_index"] == y["tri_index"]).all() and (x["uvd"].view(np.uint32) == y["uvd"].view(np.uint32)).all()
    occ = (r.occluded(batch) == r.occluded(batch, brute_force=True)).all()
    print(f"{name}: {len(batch)} rays, hit rate {x['hit'].mean():.3f}, on mesh {(x['mesh'] == 3).mean():.3f}, BVH==brute {same} anyhit {occ} (bvh {t1:.3f}s brute {t2:.3f}s)", flush=True)
s, c = r.render_accumulate(); r.reset_stats()
s, c = r.render_accumulate(); st = r.stats()
print(f"render 1920x1080 N={N}: {st['rays']} rays, device {st['device_ms']:.1f} ms, trace {st['trace_ms']:.1f} ms -> {st['rays']/st['device_ms']/1e3:.0f} Mrays/s; launches {st['kernel_launches']}", flush=True)
