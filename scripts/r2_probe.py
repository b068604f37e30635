"""Round-2 probes on one GPU (device time from CUDA events inside the library, best of 3 frames each):
  own    RT_OWN_TRI=0/1 on the cfg2 frame (own-triangle pre-test of shadow rays in k_shade)
  batch  samples_per_batch sweep on the cfg2 frame (does an L2-sized wavefront batch pay?)
  cfg4   -m 0 -N 1 -p 500000 -k 50 end to end: emission, kd build, render (BASELINE configs[3]; reference: 15 s)
usage: python scripts/r2_probe.py own batch cfg4"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import ray_tracing_engine_b200 as rt

def scene(name):
    s = rt.Scene.load(os.path.join(ROOT, "tests/golden/scenes", name + ".rtscene"))
    s.w = s.h = 420
    return s

def frames(r, n=3):
    s, c = r.render_accumulate()
    best, per = 1e9, None
    for _ in range(n):
        r.reset_stats()
        s, c = r.render_accumulate()
        st = r.stats()
        if st["device_ms"] < best:
            best, per = st["device_ms"], st
    return best, per, float(s.astype(np.float64).sum()), int(c.sum())

what = sys.argv[1:] or ["own", "batch", "cfg4"]
if "kd" in what:  # kd-tree build: host nth_element (reference tree), host canonical, device canonical
    st_scene = scene("stock")
    for photons in (50000, 500000):
        for label, flags, env in (("host reference tree", 0, None), ("host canonical", rt.RT_FLAG_KNN_EXACT, "host"),
                                  ("device canonical", rt.RT_FLAG_KNN_EXACT, None)):
            if env:
                os.environ["RT_KD_BUILD"] = env
            else:
                os.environ.pop("RT_KD_BUILD", None)
            r = rt.Renderer(st_scene, 1, 0, None, photons, 10, seed=1, flags=flags)
            best = 1e9
            for _ in range(4):
                t0 = time.perf_counter()
                r.build_photon_map()
                wall = 1e3 * (time.perf_counter() - t0)
                st = r.stats()
                best = min(best, st["kd_build_ms"])
            print(json.dumps(dict(probe="kd_build", photons=photons, stored=st["photons_stored"], builder=label,
                                  kd_build_ms=round(best, 3), build_photon_map_wall_ms=round(wall, 2),
                                  emit_ms=round(st["photon_ms"], 3))), flush=True)
            r.close()
    os.environ.pop("RT_KD_BUILD", None)
ex = scene("example")
if "own" in what:
    for v in ("0", "1"):
        os.environ["RT_OWN_TRI"] = v
        r = rt.Renderer(ex, 128, 1, seed=1)
        ms, st, chk, hits = frames(r)
        print(json.dumps(dict(probe="own_tri", RT_OWN_TRI=v, frame_ms=round(ms, 3), grays=round(st["rays"] / ms / 1e6, 3),
                              kernel_ms={k: round(x, 3) for k, x in st["kernel_ms"].items()}, checksum=chk, hits=hits)), flush=True)
        r.close()
    os.environ.pop("RT_OWN_TRI")
if "batch" in what:
    for spb in (128, 64, 32, 16, 8, 4, 2, 1):
        r = rt.Renderer(ex, 128, 1, seed=1, samples_per_batch=spb)
        t = time.perf_counter()
        ms, st, chk, hits = frames(r, 2)
        wall = (time.perf_counter() - t) / 3
        print(json.dumps(dict(probe="batch", samples_per_batch=spb, state_MB=round(spb * 176400 * 282 / 1e6), frame_ms=round(ms, 3),
                              host_wall_ms=round(1e3 * wall, 2), launches=st["kernel_launches"],
                              kernel_ms={k: round(x, 3) for k, x in st["kernel_ms"].items()}, checksum=chk)), flush=True)
        r.close()
if "cfg4" in what:
    st_scene = scene("stock")
    for it in range(3):
        t0 = time.perf_counter()
        r = rt.Renderer(st_scene, 1, 0, None, 500000, 50, seed=1)
        img = rt.Image(420, 420).fillBackground()
        r.render(img)
        wall = time.perf_counter() - t0
        st = r.stats()
        print(json.dumps(dict(probe="cfg4_e2e", iteration=it, wall_ms=round(1e3 * wall, 2), create_ms=round(st["create_ms"], 2),
                              emit_ms=round(st["photon_ms"], 3), kd_build_ms=round(st["kd_build_ms"], 2),
                              render_device_ms=round(st["device_ms"], 3), photons=st["photons_stored"],
                              photon_rays=st["photon_rays"], knn_queries=st["knn_queries"], kd_visits=st["kd_visits"],
                              kernel_ms={k: round(x, 3) for k, x in st["kernel_ms"].items()}, mean=float(img.pixels.mean()))), flush=True)
        r.close()
if "e2e" in what:  # where the host side of one user-facing cfg2 call goes (bench.py's e2e step, phase by phase)
    bg = rt.Image(420, 420).fillBackground().pixels
    acc = {}
    n_it = 12
    for it in range(n_it + 2):
        t = [time.perf_counter()]
        r = rt.Renderer(ex, 128, 1, seed=1); t.append(time.perf_counter())
        img = rt.Image(420, 420); img.pixels = bg.copy(); t.append(time.perf_counter())
        r.render(img); t.append(time.perf_counter())
        st = r.stats(); t.append(time.perf_counter())
        r.close(); t.append(time.perf_counter())
        if it < 2:
            continue
        for name, a, b in (("Renderer()", 0, 1), ("image", 1, 2), ("render", 2, 3), ("stats", 3, 4), ("close", 4, 5), ("total", 0, 5)):
            acc[name] = acc.get(name, 0.0) + 1e3 * (t[b] - t[a]) / n_it
        for key in ("create_ms", "bvh_build_ms", "device_ms"):
            acc[key] = acc.get(key, 0.0) + st[key] / n_it
    print(json.dumps(dict(probe="e2e_phases", **{k: round(v, 3) for k, v in acc.items()})), flush=True)
if "bvh" in what:  # BVH build at 11 666 triangles: host / device per-level launches / device one CTA per mesh
    for label, env in (("host", {"RT_BVH_BUILD": "host"}), ("device, per-level launches", {"RT_BVH_BUILD": "gpu", "RT_BVH_SMALL": "0"}),
                       ("device, one CTA per mesh", {"RT_BVH_BUILD": "gpu", "RT_BVH_SMALL": "1"})):
        os.environ.update(env)
        best_b, best_c = 1e9, 1e9
        for _ in range(8):
            r = rt.Renderer(ex, 1, 1, seed=1)
            st = r.stats()
            best_b, best_c = min(best_b, st["bvh_build_ms"]), min(best_c, st["create_ms"])
            launches = st["kernel_launches"]
            r.close()
        print(json.dumps(dict(probe="bvh_build", builder=label, triangles=ex.T, bvh_build_ms=round(best_b, 3),
                              create_ms=round(best_c, 3), launches=launches)), flush=True)
        for k in env:
            os.environ.pop(k)
if "sortbits" in what:  # Morton cells per axis for the binning of hit points: 32 (shared-memory counters) vs 2^bits (global)
    for photons in (0, 50000):
        for bits in ("0", "5", "6", "7"):
            os.environ["RT_SORT_BITS"] = bits
            r = rt.Renderer(ex, 128, 1, None, photons, 10, seed=1)
            if photons:
                r.build_photon_map()
            ms, st, chk, hits = frames(r, 2)
            print(json.dumps(dict(probe="sort_bits", photons=photons, RT_SORT_BITS=bits, frame_ms=round(ms, 3),
                                  kernel_ms={k: round(x, 3) for k, x in st["kernel_ms"].items()}, checksum=chk, hits=hits)), flush=True)
            r.close()
    os.environ.pop("RT_SORT_BITS")
if "tilesort" in what:  # photon k_shade: tile-local Morton ordering of the queries x global binning granularity
    for bits in os.environ.get("PROBE_BITS", "0,6").split(","):
        for ts in os.environ.get("PROBE_ROUNDS", "0,4,8,16,32").split(","):
            os.environ["RT_SORT_BITS"] = bits
            os.environ["RT_SHADE_TILE_ROUNDS"] = ts
            r = rt.Renderer(ex, 128, 1, None, 50000, 10, seed=1)
            r.build_photon_map()
            ms, st, chk, hits = frames(r, 2)
            print(json.dumps(dict(probe="tile_sort", RT_SORT_BITS=bits, RT_SHADE_TILE_ROUNDS=ts, frame_ms=round(ms, 3),
                                  kernel_ms={k: round(x, 3) for k, x in st["kernel_ms"].items() if x}, checksum=chk, hits=hits,
                                  kd_visits=st["kd_visits"])), flush=True)
            r.close()
    os.environ.pop("RT_SORT_BITS"); os.environ.pop("RT_SHADE_TILE_ROUNDS")
if "cfg5bits" in what:  # the 1.23 M-triangle scene (L2-latency-bound traversal): does finer binning of the bounce hits pay there?
    import bench
    sc5, W5, H5, N5, mode5, _, _, _ = bench.load_workload(rt, "cfg5", 0)
    for bits in ("0", "6", "7"):
        os.environ["RT_SORT_BITS"] = bits
        r = rt.Renderer(sc5, 16, mode5, seed=1)
        ms, st, chk, hits = frames(r, 2)
        print(json.dumps(dict(probe="cfg5_sort_bits", RT_SORT_BITS=bits, N=16, frame_ms=round(ms, 3), grays=round(st["rays"] / ms / 1e6, 3),
                              kernel_ms={k: round(x, 3) for k, x in st["kernel_ms"].items() if x}, checksum=chk, hits=hits)), flush=True)
        r.close()
    os.environ.pop("RT_SORT_BITS")
if "envab" in what:  # generic A/B of an environment switch on the cfg2 frame: PROBE_ENV="RT_X=0,1"
    var, vals = os.environ["PROBE_ENV"].split("=")
    for v in vals.split(","):
        os.environ[var] = v
        r = rt.Renderer(ex, 128, 1, seed=1)
        ms, st, chk, hits = frames(r)
        print(json.dumps(dict(probe="env_ab", env=f"{var}={v}", frame_ms=round(ms, 3), grays=round(st["rays"] / ms / 1e6, 3),
                              kernel_ms={k: round(x, 3) for k, x in st["kernel_ms"].items() if x}, checksum=chk, hits=hits)), flush=True)
        r.close()
    os.environ.pop(var)
