"""gen_golden.py -- regenerate tests/golden/ from the UNMODIFIED reference (oracle/_ref/libref_cb.so).

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py [--only scenes,rng,...] [--heavy]

Every file written here is an OUTPUT OF THE REFERENCE'S OWN CODE (scene assembly, loadOFF, rayTrace,
evaluateColorResponse, the samplers, PhotonMap, kdtree::knearest, calculateColor*), driven through
oracle/ref_harness.cpp with the shared counter-based RNG engine.  The GPU box has no /root/reference,
so the parity tests there compare against these vectors (and against the CPU restatement, which the
CPU-side tests pin against the same library bit for bit).
"""
from __future__ import annotations

import argparse
import hashlib
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[0] = ROOT  # (the script dir would shadow the oracle package with oracle.py)
from oracle import oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MESHES = "/root/reference/meshes"
SEED = 1


def log(*a):
    print("[gen_golden]", *a, flush=True)


def save(name, **arrays):
    path = os.path.join(GOLD, name)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **arrays)
    log("wrote", name, f"{os.path.getsize(path) / 1024:.0f} KiB")


def scenes():
    ref = O.RefOracle()
    os.makedirs(os.path.join(GOLD, "scenes"), exist_ok=True)
    for name, off in (("stock", None), ("lowres", f"{MESHES}/example_low_res.off"), ("example", f"{MESHES}/example.off")):
        flat = ref.create_scene(420, 420, custom_off=off)
        flat.save(os.path.join(GOLD, "scenes", f"{name}.rtscene"))
        log("scene", name, "V,T,M,L =", flat.V, flat.T, flat.M, flat.L)
    # non-square camera (the CLI default 380x270) for the host-side camera test
    flat = ref.create_scene(380, 270)
    save("camera_380x270.npz", cam=flat.cam)
    out = {}
    for key, f in (("cube_tri", "cube_tri.off"), ("cube_tri2", "cube_tri2.off"), ("example_low_res", "example_low_res.off")):
        pos, nrm, tri = ref.load_off(f"{MESHES}/{f}")
        out[key + "_pos"], out[key + "_tri"] = pos, tri
    save("meshes_loaded.npz", **out)
    save("background.npz", rows_420=ref.background(420, 420)[:, 0, :], rows_270=ref.background(380, 270)[:, 0, :])
    # the reference's OFF loader on our own tiny fixture (quads + comment + polygon fan)
    pos, nrm, tri = ref.load_off(os.path.join(GOLD, "fixture_mixed.off"))
    save("fixture_mixed_loaded.npz", pos=pos, nrm=nrm, tri=tri)


def rng():
    ref = O.RefOracle()
    out = {}
    for i, idx in enumerate((0, 5, 123456789012)):
        out[f"words_{i}"] = ref.rng_words(SEED, O.DOMAIN_PIXEL, idx, 64)
        out[f"uf_{i}"] = ref.rng_uniform_float(SEED, O.DOMAIN_PIXEL, idx, 256, -0.01, 0.01)
        out[f"ud_{i}"] = ref.rng_uniform_double(SEED, O.DOMAIN_PHOTON, idx, 256, 0.0, 1.0000000278275352)
        out[f"idx_{i}"] = np.array([idx], np.uint64)
    save("rng.npz", **out)


def sampling():
    ref = O.RefOracle()
    ref.create_scene(420, 420)
    g = np.random.default_rng(11)
    normals = g.normal(size=(4096, 3)).astype(np.float32)
    out = dict(normals=normals,
               hsphere=ref.hsphere(SEED, O.DOMAIN_PHOTON, 1000, normals),
               jitter_0_1=ref.jitter(SEED, O.DOMAIN_PIXEL, 0, 4096, 0, 1),
               jitter_77_128=ref.jitter(SEED, O.DOMAIN_PIXEL, 0, 4096, 77, 128),
               jitter_1000_1024=ref.jitter(SEED, O.DOMAIN_PIXEL, 0, 4096, 1000, 1024))
    for l in range(3):
        out[f"light_{l}"] = ref.light_sample(l, SEED, O.DOMAIN_PIXEL, 0, 4096)
    pts = g.uniform(-2, 2, (2048, 3)).astype(np.float32)
    out["pts"] = pts
    for l in range(3):
        out[f"light_eval_{l}"] = ref.light_eval(l, pts)
    xy = np.stack(np.meshgrid(np.arange(0, 420, 7), np.arange(0, 420, 7)), -1).reshape(-1, 2).astype(np.int32)
    shift = g.uniform(0, 1, (len(xy), 2)).astype(np.float32)
    out["cam_xy"], out["cam_shift"], out["cam_rays"] = xy, shift, ref.camera_rays(xy, shift)
    save("sampling.npz", **out)


def bsdf():
    ref = O.RefOracle()
    flat = ref.create_scene(420, 420)
    g = np.random.default_rng(12)
    v = g.normal(size=(8192, 9)).astype(np.float32)
    # the two hand-checked cases of SURVEY.md 8a-B
    v[0] = [0, 1, 0, .3, 1, .2, -.2, 1, .1]
    out = dict(inputs=v, mats=flat.mats)
    for m in range(flat.M):
        out[f"bsdf_{m}"] = ref.bsdf(flat.mats[m], v)
    save("bsdf.npz", **out)


def make_rays(ref, flat, n_random, seed):
    """Primary rays on a grid, random rays, and shadow/bounce-like rays leaving hit points."""
    g = np.random.default_rng(seed)
    xy = np.stack(np.meshgrid(np.arange(0, 420, 5), np.arange(0, 420, 5)), -1).reshape(-1, 2).astype(np.int32)
    prim = ref.camera_rays(xy, g.uniform(0, 1, (len(xy), 2)).astype(np.float32))
    rnd = np.concatenate([g.uniform(-1.5, 1.5, (n_random, 3)), g.normal(size=(n_random, 3))], 1).astype(np.float32)
    h = ref.trace(prim)
    hitp = prim[:, :3] + prim[:, 3:] * h["uvd"][:, 2:3]
    ok = h["hit"] == 1
    # secondary rays start (almost) ON a surface, like the reference's shadow and bounce rays
    lights = flat.lights[:, :3]
    sec = []
    for l in range(len(lights)):
        sec.append(np.concatenate([hitp[ok], lights[l][None] - hitp[ok]], 1))
    sec.append(np.concatenate([hitp[ok], g.normal(size=(ok.sum(), 3))], 1))
    return np.concatenate([prim, rnd] + sec).astype(np.float32)


def trace():
    ref = O.RefOracle()
    for name, off, nrand in (("stock", None, 20000), ("lowres", f"{MESHES}/example_low_res.off", 6000),
                             ("example", f"{MESHES}/example.off", 3000)):
        flat = ref.create_scene(420, 420, custom_off=off)
        rays = make_rays(ref, flat, nrand, 21)
        if name == "example":
            rays = rays[::3]
        t = time.time()
        h = ref.trace(rays)
        log("trace", name, len(rays), "rays", f"{time.time() - t:.1f}s", "hit rate", h["hit"].mean())
        save(f"trace_{name}.npz", rays=rays, hit=h["hit"].astype(np.int8), mesh=h["mesh"].astype(np.int8),
             tri3=h["tri3"], uvd=h["uvd"])


def to8(img):
    """Image::savePPM quantisation (Image.cpp:31-38): unsigned(255.f * v)."""
    return (np.float32(255.0) * img.astype(np.float32)).astype(np.uint32).astype(np.uint8)


def render_case(ref, name, N, mode, window=None, photons=0, k=0, keep_float=True, want_samples=False):
    pm = ref.photon_map_create(photons, SEED) if photons else None
    t = time.time()
    r = ref.render(N, mode, SEED, num_photons=photons, k=k, photon_map=pm, window=window, want_samples=want_samples)
    w, h = ref.flat.w, ref.flat.h
    out = dict(N=np.array([N]), mode=np.array([mode]), photons=np.array([photons]), k=np.array([k]),
               window=np.array(window if window else (0, 0, w, h)), counter=r["counter"].astype(np.int16))
    if keep_float:
        out["sum_rgb"] = r["sum_rgb"]
    if window is None:
        img = ref.composite(N, r["sum_rgb"], r["counter"], ref.background(w, h))
        out["image8"] = to8(img)
    if want_samples:
        out["samples"] = r["samples"]
        out["found"] = r["found"]
    if pm is not None:
        plist, hist = pm.get()
        out["photon_count"] = np.array([len(plist)])
        out["depth_hist"] = hist
    log("render", name, f"{time.time() - t:.1f}s")
    save(f"render_{name}.npz", **out)


def render_light():
    ref = O.RefOracle()
    ref.create_scene(420, 420)
    render_case(ref, "stock_m0_N1", 1, 0)
    render_case(ref, "stock_m1_N4_win", 4, 1, window=(100, 150, 228, 214), want_samples=True)
    render_case(ref, "stock_m0_p3000_k10_win", 1, 0, window=(100, 150, 228, 214), photons=3000, k=10,
                want_samples=True)
    render_case(ref, "stock_m1_p3000_k5_N2_win", 2, 1, window=(100, 150, 228, 214), photons=3000, k=5,
                want_samples=True)
    ref.create_scene(420, 420, custom_off=f"{MESHES}/example_low_res.off")
    render_case(ref, "lowres_m0_N1_win", 1, 0, window=(120, 100, 248, 228), want_samples=True)
    render_case(ref, "lowres_m1_N2_win", 2, 1, window=(150, 130, 214, 194), want_samples=True)


def render_heavy():
    """Minutes of reference CPU time each."""
    ref = O.RefOracle()
    ref.create_scene(420, 420)
    render_case(ref, "stock_m1_N128", 128, 1, keep_float=True)
    render_case(ref, "stock_m0_p50000_k10_N1", 1, 0, photons=50000, k=10, keep_float=False)
    ref.create_scene(420, 420, custom_off=f"{MESHES}/example.off")
    render_case(ref, "example_m1_N2_win", 2, 1, window=(170, 150, 234, 214), want_samples=True)
    render_case(ref, "example_m0_N1", 1, 0, keep_float=False)


def photons():
    ref = O.RefOracle()
    ref.create_scene(420, 420)
    pm = ref.photon_map_create(3000, SEED)
    plist, hist = pm.get()
    nodes, left, right, root = pm.layout()
    g = np.random.default_rng(31)
    q = np.concatenate([g.uniform(-1.5, 1.5, (1500, 3)), plist[g.integers(0, len(plist), 500), :3] +
                        g.normal(scale=1e-3, size=(500, 3))]).astype(np.float32)
    q[-1] = plist[7, :3]  # an exact hit: m_bestdist == 0 path (kdtree.h:101)
    out = dict(list=plist, hist=hist, nodes=nodes, left=left, right=right, root=np.array([root]), queries=q)
    for k in (1, 5, 10, 50):
        res, visited = pm.knn(q, k)
        out[f"knn_{k}"] = res[:, :, :3]  # positions identify the photons
        out[f"visited_{k}"] = visited.astype(np.int32)
    # depth histogram of the two BASELINE photon counts (statistical golden for emission)
    for n in (50000,):
        pm2 = ref.photon_map_create(n, SEED)
        l2, h2 = pm2.get()
        out[f"hist_{n}"] = h2
        out[f"count_{n}"] = np.array([len(l2)])
    save("photons.npz", **out)


def pointcloud():
    """PhotonMap::saveToPCD (PhotonMap.h:59-84): the reference's own writer on the first 64 golden photons plus a few
    values that exercise the stream formatting (exponents, negative zero, integers)."""
    ref = O.RefOracle()
    plist = np.load(os.path.join(GOLD, "photons.npz"))["list"][:64].copy()
    plist[0, :6] = [1.0, -0.0, 1e-7, 123456.789, -2.5e10, 0.1]
    np.save(os.path.join(GOLD, "pointcloud_input.npy"), plist)
    ref.save_pcd(plist, os.path.join(GOLD, "pointcloud_golden.pcd"))
    log("wrote pointcloud_golden.pcd", os.path.getsize(os.path.join(GOLD, "pointcloud_golden.pcd")), "bytes")


def stock_binary():
    """md5 of the unmodified reference program's own output (stock RNG) -- guards oracle drift."""
    build = os.path.join(HERE, "_ref", "build")
    os.makedirs(build, exist_ok=True)
    subprocess.run([O.REF_BIN, "-width", "420", "-height", "420", "-m", "0", "-N", "1"], cwd=build, check=True,
                   stdout=subprocess.DEVNULL)
    md5 = hashlib.md5(open(os.path.join(build, "output.ppm"), "rb").read()).hexdigest()
    log("stock binary -m 0 -N 1 420x420 md5", md5)
    with open(os.path.join(GOLD, "stock_binary_md5.txt"), "w") as f:
        f.write(f"{md5}  RayTracer -width 420 -height 420 -m 0 -N 1 (unmodified reference, g++ 13.3 -O3 "
                f"-ffp-contract=off, stock minstd_rand0 engine)\n")


# ------------------------------------------------------------------------------------------------
# Round 2: the configs the north_star target names, at their own sample count (N = 128), and an
# INDEPENDENT converged mean from the reference's own serial minstd_rand0 engine.
# ------------------------------------------------------------------------------------------------
def sample_hash16(samples):
    """16-bit fingerprint of every clamped sample colour (3 binary32 words): lets a test count bit-identical
    samples without storing 12 bytes per sample.  Restated in tests/test_gpu_parity.py."""
    b = np.ascontiguousarray(samples, np.float32).view(np.uint32).astype(np.uint64)
    h = (b[..., 0] * np.uint64(0x9E3779B1) ^ b[..., 1]) * np.uint64(0x85EBCA77) ^ b[..., 2]
    h ^= h >> np.uint64(29)
    h = (h * np.uint64(0xC2B2AE3D)) & np.uint64(0xFFFFFFFFFFFF)
    return ((h >> np.uint64(24)) & np.uint64(0xFFFF)).astype(np.uint16)


def _band_worker(args):
    off, N, mode, photons, k, plist, window, rows = args
    ref = O.RefOracle()
    ref.create_scene(420, 420, custom_off=off)
    pm = ref.photon_map_from_list(plist) if plist is not None else None
    x0, _, x1, _ = window
    r = ref.render(N, mode, SEED, num_photons=photons, k=k, photon_map=pm, window=(x0, rows[0], x1, rows[1]),
                   want_samples=True)
    return dict(sum_rgb=r["sum_rgb"], counter=r["counter"], hash16=sample_hash16(r["samples"]),
                found=r["found"].astype(np.int8))


def render_window_parallel(off, N, mode, window, photons=0, k=0, plist=None, procs=8):
    import multiprocessing as mp
    x0, y0, x1, y1 = window
    rows = y1 - y0
    bands = [(y0 + rows * i // procs, y0 + rows * (i + 1) // procs) for i in range(procs)]
    jobs = [(off, N, mode, photons, k, plist, window, b) for b in bands if b[1] > b[0]]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        parts = pool.map(_band_worker, jobs)
    return dict(sum_rgb=np.concatenate([p["sum_rgb"] for p in parts], 0),
                counter=np.concatenate([p["counter"] for p in parts], 0),
                hash16=np.concatenate([p["hash16"] for p in parts], 1),
                found=np.concatenate([p["found"] for p in parts], 1))


def headline_windows():
    """BASELINE configs[1] and [2] at their own N: example.off scene, -m 1 -N 128, a 64x64 window, every sample
    (sum, counter, primary-hit flags and a 16-bit fingerprint per sample), without and with a 50 000-photon map
    emitted by the reference (the shared list is committed with the golden)."""
    off = f"{MESHES}/example.off"
    win = (170, 150, 234, 214)
    t = time.time()
    r = render_window_parallel(off, 128, 1, win)
    log("headline cfg2 window", f"{time.time() - t:.0f}s")
    save("render_example_m1_N128_win.npz", window=np.array(win), N=np.array([128]), sum_rgb=r["sum_rgb"],
         counter=r["counter"].astype(np.int16), hash16=r["hash16"], found=np.packbits(r["found"].astype(bool)))
    ref = O.RefOracle()
    ref.create_scene(420, 420, custom_off=off)
    t = time.time()
    pm = ref.photon_map_create(50000, SEED)
    plist, hist = pm.get()
    log("example-scene photon emission", len(plist), "stored", f"{time.time() - t:.0f}s")
    t = time.time()
    r = render_window_parallel(off, 128, 1, win, photons=50000, k=10, plist=plist)
    log("headline cfg3 window", f"{time.time() - t:.0f}s")
    save("render_example_m1_N128_p50000_k10_win.npz", window=np.array(win), N=np.array([128]), photons=plist,
         depth_hist=hist, sum_rgb=r["sum_rgb"], counter=r["counter"].astype(np.int16), hash16=r["hash16"],
         found=np.packbits(r["found"].astype(bool)))


def _stock_mean_worker(args):
    seed, W, H, N = args
    ref = O.RefOracle(stock_rng=True)
    ref.create_scene(W, H)
    ref.lib.ref_reseed(seed)
    acc = np.zeros((H, W, 3), np.float64)
    acc2 = np.zeros((H, W, 3), np.float64)
    cnt = np.zeros((H, W), np.int64)
    chunk = 64
    for s0 in range(0, N, chunk):  # sample ranges in order: the serial engine is consumed exactly as one N-sample run
        r = ref.render(N, 1, 0, samples=(s0, min(N, s0 + chunk)), want_samples=True)
        smp = r["samples"].astype(np.float64)
        acc += smp.sum(0)
        acc2 += (smp * smp).sum(0)
        cnt += r["counter"]
    return acc, acc2, cnt


def converged_mean():
    """SURVEY.md section 4 test 5: the reference's converged mean from its OWN engine (std::default_random_engine =
    minstd_rand0 consumed serially, LightSource.h:6) -- statistically independent of the counter-based streams the
    GPU and the shared-RNG oracle use.  Stock scene, 105x105, -m 1, N = 2048, 3 seeds: per-pixel mean and per-pixel
    sample variance (float64 accumulation of the clamped sample colours)."""
    import multiprocessing as mp
    W = H = 105
    N, seeds = 2048, (1, 2, 3)
    t = time.time()
    with mp.get_context("fork").Pool(len(seeds)) as pool:
        parts = pool.map(_stock_mean_worker, [(s, W, H, N) for s in seeds])
    n = N * len(seeds)
    mean_seed = np.stack([p[0] / N for p in parts])
    acc = sum(p[0] for p in parts)
    acc2 = sum(p[1] for p in parts)
    mean = acc / n
    var = (acc2 / n - mean * mean) * n / (n - 1)
    hitfrac = sum(p[2] for p in parts) / n
    seed_rmse = [float(np.sqrt(np.mean((mean_seed[i] - mean_seed[j]) ** 2))) for i, j in ((0, 1), (0, 2), (1, 2))]
    log("converged mean", f"{time.time() - t:.0f}s", "seed-to-seed RMSE of the N=2048 means", seed_rmse,
        "sqrt(2*mean var/N) =", float(np.sqrt(2 * var.mean() / N)))
    save("converged_stock_m1_105.npz", W=np.array([W]), N=np.array([N]), seeds=np.array(seeds),
         mean=mean.astype(np.float32), var=var.astype(np.float32), hit_fraction=hitfrac.astype(np.float32),
         mean_per_seed=mean_seed.astype(np.float32), seed_rmse=np.array(seed_rmse))


def light_variants():
    """Scenes with 1 and 5 light sources (Scene::lightsources() may hold any number: Scene.h:14-26, Renderer.cpp:49,
    PhotonMap.h:19-24) and a k = 100 gather (kdtree::knearest accepts any k <= nodes, kdtree.h:180-183).  The light
    bases come from the reference's own LightSource constructor; renders are the reference's, shared streams."""
    ref = O.RefOracle()
    base = ref.create_scene(420, 420)
    extra = np.array([[0.9, 1.2, 0.4, 1.0, 0.9, 0.7, 0.0, 0.0, 0.0, 0.85, 0.05],
                      [-0.7, 0.9, 1.2, 0.6, 0.8, 1.0, 0.1, 0.0, -0.2, 0.85, 0.02]], np.float32)
    win = (100, 150, 228, 214)
    for name, ctor in (("stock_1light", base.lights_ctor[2:3]), ("stock_5lights", np.concatenate([base.lights_ctor, extra]))):
        flat = O.FlatScene(base.pos, base.nrm, base.tri, base.mesh_tri_off, base.mesh_vtx_off, base.mats,
                           np.zeros((len(ctor), 21), np.float32), ctor, base.cam, 420, 420)
        ref.set_scene(flat)
        flat = ref.flatten(420, 420)  # bases as the reference's constructor computed them (LightSource.h:29-32)
        flat.save(os.path.join(GOLD, "scenes", f"{name}.rtscene"))
        ref.set_scene(flat)
        out = {}
        for mode, N in ((0, 1), (1, 3)):
            r = ref.render(N, mode, SEED, window=win, want_samples=True)
            out[f"samples_m{mode}"], out[f"found_m{mode}"] = r["samples"], r["found"]
        pm = ref.photon_map_create(3000, SEED)
        plist, hist = pm.get()
        out["photons"], out["photon_hist"] = plist, hist
        save(f"render_{name}_win.npz", window=np.array(win), **out)
    # k = 100 on the stock scene's 3000-photon list: query results and a gather window
    ref.create_scene(420, 420)
    pm = ref.photon_map_create(3000, SEED)
    plist, _ = pm.get()
    g = np.random.default_rng(77)
    q = np.concatenate([g.uniform(-1.5, 1.5, (300, 3)), plist[g.integers(0, len(plist), 100), :3]]).astype(np.float32)
    out = dict(queries=q)
    for k in (65, 100, 300):
        res, _ = pm.knn(q, k)
        out[f"knn_{k}"] = res[:, :, :3]
    r = ref.render(1, 0, SEED, num_photons=3000, k=100, photon_map=pm, window=(100, 150, 164, 182), want_samples=True)
    out["window"], out["samples_m0_k100"], out["found_m0_k100"] = np.array((100, 150, 164, 182)), r["samples"], r["found"]
    save("knn_large_k.npz", **out)


ALL = dict(scenes=scenes, rng=rng, sampling=sampling, bsdf=bsdf, trace=trace, photons=photons,
           render_light=render_light, stock_binary=stock_binary)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--heavy", action="store_true")
    a = ap.parse_args()
    O.build(ref=True)
    names = [n for n in a.only.split(",") if n] or list(ALL)
    for n in names:
        (ALL | dict(render_heavy=render_heavy, headline_windows=headline_windows, converged_mean=converged_mean,
                                                    light_variants=light_variants, pointcloud=pointcloud))[n]()
    if a.heavy and "render_heavy" not in names:
        render_heavy()
