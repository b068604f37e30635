"""GPU edge cases of the render path, each against the CPU oracle or a property the reference guarantees:
empty / tiny / degenerate geometry, N = 0, 1x1 and ragged image sizes, fewer than three lights, more shards than
tiles, the largest k, scenes with more meshes than the root list holds."""
import numpy as np
import pytest

from conftest import scene_path

pytestmark = pytest.mark.gpu


def beq(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and bool((a.view(np.uint32) == b.view(np.uint32)).all())


@pytest.fixture(scope="module")
def rt():
    import ray_tracing_engine_b200 as m
    if m.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need the B200 (there is no CPU fallback)")
    return m


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def sub_scene(rt, scene, keep_meshes, lights=None):
    """The stock scene restricted to some meshes (vertices kept as they are, triangles re-offset per mesh)."""
    tri, off = [], [0]
    for m in keep_meshes:
        a, b = scene.mesh_tri_off[m], scene.mesh_tri_off[m + 1]
        tri.append(scene.tri.reshape(-1, 3)[a:b])
        off.append(off[-1] + (b - a))
    tri = np.concatenate(tri) if tri else np.zeros((0, 3), np.int32)
    vtx = np.zeros(len(keep_meshes) + 1, np.int32)  # global vertex ids are kept, so per-mesh vertex offsets are 0
    L = scene.lights.reshape(-1, 21)
    L = L if lights is None else L[:lights]
    return rt.Scene(scene.pos, scene.nrm, tri, np.int32(off), vtx, scene.mats.reshape(-1, 8)[list(keep_meshes)],
                    L, scene.cam, scene.w, scene.h, None)


def test_empty_scene_renders_the_background(rt):
    """No triangles at all: every primary ray misses, counter stays 0, and the composite is the reference's
    `0/N + background*(N-0)/N` (Renderer.cpp:262-265) -- which is not bit-for-bit the background in binary32."""
    stock = rt.Scene.load(scene_path("stock"))
    scene = sub_scene(rt, stock, [])
    r = rt.Renderer(scene, 3, 1, seed=1, width=40, height=24)
    bg = rt.Image(40, 24).fillBackground()
    s, c = r.render_accumulate()
    assert not s.any() and not c.any()
    img = r.render(rt.Image(40, 24).fillBackground())
    assert beq(img.pixels, rt.Renderer.composite(3, s, c, bg.pixels))
    assert np.abs(img.pixels - bg.pixels).max() < 1e-6
    h = r.rayTrace(np.float32([[0, 0, 3, 0, 0, -1]]))
    assert h["hit"][0] == 0 and r.occluded(np.float32([[0, 0, 3, 0, 0, -1]]))[0] == 0


def test_single_triangle_and_single_mesh_scenes(rt):
    """T = 1 (a root with one leaf) and one 2-triangle mesh (no top-level join, no root list): BVH == brute force."""
    stock = rt.Scene.load(scene_path("stock"))
    g = np.random.default_rng(2)
    rays = np.concatenate([np.tile(np.float32([0.3, 0.6, 2.3]), (4000, 1)),
                           (g.normal(size=(4000, 3)) * [0.4, 0.4, 0.1] + [-0.1, -0.3, -1]).astype(np.float32)], 1)
    one_mesh = sub_scene(rt, stock, [0])
    tri1 = rt.Scene(stock.pos, stock.nrm, one_mesh.tri.reshape(-1, 3)[:1], np.int32([0, 1]), np.int32([0, 0]),
                    stock.mats.reshape(-1, 8)[:1], stock.lights, stock.cam, stock.w, stock.h, None)
    for scene in (tri1, one_mesh, sub_scene(rt, stock, [0, 3])):
        r = rt.Renderer(scene, 1, 0, seed=1, width=32, height=32)
        a, b = r.rayTrace(rays), r.rayTrace(rays, brute_force=True)
        assert (a["tri_index"] == b["tri_index"]).all() and beq(a["uvd"], b["uvd"])
        assert (r.occluded(rays) == r.occluded(rays, brute_force=True)).all()
        s1, c1 = r.render_accumulate()
        r.set(flags=1)  # RT_FLAG_BRUTE_FORCE: the reference's own O(T) scan
        s2, c2 = r.render_accumulate()
        assert beq(s1, s2) and (c1 == c2).all()


def test_degenerate_triangles_are_never_hit(rt, O):
    """Zero-area and needle triangles: |det| < 1e-6 rejects them in Ray.cpp:14 -- same hits as the CPU restatement."""
    stock = rt.Scene.load(scene_path("stock"))
    pos = stock.pos.reshape(-1, 3).copy()
    tri = stock.tri.reshape(-1, 3).copy()
    t0 = stock.mesh_tri_off[-2]  # first triangle of the last mesh
    tri[t0] = [tri[t0][0], tri[t0][0], tri[t0][1]]          # two equal vertices: zero area
    tri[t0 + 1] = [tri[t0 + 1][0], tri[t0 + 1][1], tri[t0 + 1][1]]
    scene = rt.Scene(pos, stock.nrm, tri, stock.mesh_tri_off, stock.mesh_vtx_off, stock.mats, stock.lights, stock.cam,
                     stock.w, stock.h, stock.lights_ctor)
    flat = O.FlatScene(scene.pos, scene.nrm, scene.tri, scene.mesh_tri_off, scene.mesh_vtx_off, scene.mats, scene.lights,
                       scene.lights_ctor, scene.cam, 48, 36)
    port = O.PortOracle(flat)
    want = port.render(2, 0, 7, want_samples=True)
    r = rt.Renderer(scene, 2, 0, seed=7, width=48, height=36)
    rgb, found = r.render_samples()
    assert (found == want["found"]).all() and beq(rgb, want["samples"])
    g = np.random.default_rng(3)
    rays = np.concatenate([g.uniform(-1, 1, (20000, 3)), g.normal(size=(20000, 3))], 1).astype(np.float32)
    a = r.rayTrace(rays)
    assert not np.isin(a["tri_index"], [t0, t0 + 1]).any()
    assert (a["tri_index"] == r.rayTrace(rays, brute_force=True)["tri_index"]).all()


def test_zero_samples_and_tiny_ragged_images(rt, O):
    """N = 0 gives the reference's black frame (Renderer.cpp:208,219,271: the loop does not run and `image = saveImage`,
    a zero-initialised Image); 1x1, 1xH, Wx1 and odd sizes match the CPU oracle."""
    scene = rt.Scene.load(scene_path("stock"))
    r0 = rt.Renderer(scene, 0, 1, seed=1, width=16, height=8)
    img = rt.Image(16, 8).fillBackground()
    assert not r0.render(img).pixels.any()
    for (w, h) in ((1, 1), (1, 7), (9, 1), (17, 13), (33, 5)):
        flat = O.FlatScene.load(scene_path("stock"))
        flat.w, flat.h = w, h
        port = O.PortOracle(flat)
        want = port.render(2, 0, 11, want_samples=True)
        r = rt.Renderer(scene, 2, 0, seed=11, width=w, height=h)
        rgb, found = r.render_samples()
        assert (found == want["found"]).all() and beq(rgb, want["samples"]), (w, h)
        s, c = r.render_accumulate()
        assert beq(s, want["sum_rgb"]) and (c == want["counter"]).all()


def test_fewer_than_three_lights(rt, O):
    """The shadow-ray queue is laid out for three lights; one or two lights must give the oracle's colours."""
    stock = rt.Scene.load(scene_path("stock"))
    for nl in (1, 2):
        scene = rt.Scene(stock.pos, stock.nrm, stock.tri, stock.mesh_tri_off, stock.mesh_vtx_off, stock.mats,
                         stock.lights.reshape(-1, 21)[:nl], stock.cam, stock.w, stock.h,
                         None if stock.lights_ctor is None else stock.lights_ctor.reshape(-1, 11)[:nl])
        flat = O.FlatScene(scene.pos, scene.nrm, scene.tri, scene.mesh_tri_off, scene.mesh_vtx_off, scene.mats,
                           scene.lights, scene.lights_ctor, scene.cam, 40, 30)
        want = O.PortOracle(flat).render(2, 0, 5, want_samples=True)
        rgb, found = rt.Renderer(scene, 2, 0, seed=5, width=40, height=30).render_samples()
        assert (found == want["found"]).all() and beq(rgb, want["samples"]), nl


def test_more_shards_than_tiles(rt):
    """64 ranks on a 48x32 image with 16x16 tiles: ranks 6..63 own nothing, the others sum to the 1-GPU frame."""
    scene = rt.Scene.load(scene_path("stock"))
    full_s, full_c = rt.Renderer(scene, 2, 1, seed=9, width=48, height=32).render_accumulate()
    acc_s, acc_c = np.zeros_like(full_s), np.zeros_like(full_c)
    for rank in (0, 1, 2, 3, 4, 5, 6, 40, 63):
        s, c = rt.Renderer(scene, 2, 1, seed=9, width=48, height=32, shard_rank=rank, shard_count=64).render_accumulate()
        if rank >= 6:
            assert not s.any() and not c.any()
        acc_s += s
        acc_c += c
    assert beq(acc_s, full_s) and (acc_c == full_c).all()


def test_largest_k_and_k_equal_to_photon_count(rt, O, gold):
    """k = 64 (the largest k whose candidates live in shared memory) and k == number of photons (every photon is a
    neighbour): order-identical to the oracle.  Larger k: tests/test_gpu_round2.py."""
    g = gold("photons.npz")
    scene = rt.Scene.load(scene_path("stock"))
    port = O.PortOracle(O.FlatScene.load(scene_path("stock")))
    for plist, k in ((g["list"], 64), (g["list"][:23], 23)):
        r = rt.Renderer(scene, 1, 0, None, 3000, k, seed=1)
        r.set_photons(plist)
        pm = port.photon_map_from_list(plist)
        q = g["queries"][:800]
        out, visited, oidx = pm.knn(q, k, want_index=True)
        assert (r.knearest(q, k) == oidx).all(), k
    with pytest.raises(rt.RtError):
        rt.Renderer(scene, 1, 0, None, 3000, rt._capi.RT_MAX_K + 1, seed=1)


def test_more_meshes_than_the_root_list_holds(rt):
    """20 meshes (> kMaxRoots = 16): traversal falls back to the top-level tree; hits still equal the O(T) scan,
    and mesh order still breaks exact ties (RayTracer.h:40: the first mesh in scene order wins)."""
    stock = rt.Scene.load(scene_path("stock"))
    T = stock.tri.reshape(-1, 3)
    wall = T[stock.mesh_tri_off[0]:stock.mesh_tri_off[1]]
    reps = 20
    tri = np.concatenate([wall] * reps + [T])          # 20 coincident copies of the first mesh, then the scene
    off = np.concatenate([np.arange(reps) * len(wall), reps * len(wall) + stock.mesh_tri_off]).astype(np.int32)
    mats = np.concatenate([np.tile(stock.mats.reshape(-1, 8)[:1], (reps, 1)), stock.mats.reshape(-1, 8)])
    scene = rt.Scene(stock.pos, stock.nrm, tri, off, np.zeros(len(off), np.int32), mats, stock.lights, stock.cam,
                     stock.w, stock.h, None)
    r = rt.Renderer(scene, 1, 0, seed=1, width=32, height=32)
    g = np.random.default_rng(4)
    rays = np.concatenate([np.tile(np.float32([0.3, 0.6, 2.3]), (5000, 1)),
                           (g.normal(size=(5000, 3)) * [0.5, 0.5, 0.1] + [-0.1, -0.3, -1]).astype(np.float32)], 1)
    a, b = r.rayTrace(rays), r.rayTrace(rays, brute_force=True)
    assert (a["tri_index"] == b["tri_index"]).all() and beq(a["uvd"], b["uvd"])
    on_copies = a["tri_index"][(a["tri_index"] >= 0) & (a["tri_index"] < (reps + 1) * len(wall))]
    assert len(on_copies) > 0 and (on_copies < len(wall)).all(), "exact ties must go to the first copy"


def test_million_triangle_scene_bvh_equals_bruteforce(rt, tmp_path):
    """BASELINE config 5's scene (example_low_res.off subdivided 5x, 1 228 822 triangles; BVH built on the device,
    25 levels): nearest and any-hit through the BVH return exactly what the O(T) scan returns, for camera rays and
    for rays leaving the surfaces; one 1080p sample pass is reproducible and hits the frame everywhere."""
    import os, sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from bench import materialize_meshes
    mdir = materialize_meshes(str(tmp_path / "meshes"))
    scene = rt.Scene.build(1920, 1080, mdir, os.path.join(mdir, "example_low_res.off"), 5)
    assert scene.T == 1200 * 4 ** 5 + 22
    r = rt.Renderer(scene, 1, 1, seed=1)
    st = r.stats()
    assert st["bvh_nodes"] == scene.T - 1 and st["bvh_depth"] >= 22
    g = np.random.default_rng(3)
    n = 20000
    o = np.tile(np.float32([0.3, 0.6, 2.3]), (n, 1))
    d = (g.normal(size=(n, 3)) * [0.25, 0.25, 0.1] + [-0.12, -0.25, -1]).astype(np.float32)
    rays = np.concatenate([o, d], 1).astype(np.float32)
    a = r.rayTrace(rays)
    ok = a["hit"] == 1
    assert ok.mean() > 0.99 and (a["mesh"][ok] == 3).mean() > 0.1   # a good share of the rays lands on the big mesh
    p = (o + d * a["uvd"][:, 2:3])[ok]
    rays2 = np.concatenate([p, g.normal(size=p.shape).astype(np.float32)], 1)
    for batch in (rays, rays2):
        x, y = r.rayTrace(batch), r.rayTrace(batch, brute_force=True)
        assert (x["tri_index"] == y["tri_index"]).all() and beq(x["uvd"], y["uvd"])
        assert (r.occluded(batch) == r.occluded(batch, brute_force=True)).all()
    s1, c1 = r.render_accumulate()
    s2, c2 = r.render_accumulate()
    assert beq(s1, s2) and (c1 == c2).all() and c1.min() == 1


def test_out_of_memory_retry_renders_the_same_frame(rt, monkeypatch):
    """When the default 32 M-path wavefront allocation fails, the batch shrinks to half of the free HBM and the
    render is retried (csrc/capi.cu).  RT_TEST_OOM_ONCE injects the failure and pretends 64 MB are free: the frame is
    rendered in many small batches and must be bit-identical (samples are accumulated in index order)."""
    scene = rt.Scene.load(scene_path("stock"))
    r0 = rt.Renderer(scene, 6, 1, seed=8, width=160, height=120)
    want_s, want_c = r0.render_accumulate()
    one_batch = r0.stats()["kernel_launches"]
    monkeypatch.setenv("RT_TEST_OOM_ONCE", "1")
    r = rt.Renderer(scene, 6, 1, seed=8, width=160, height=120)
    s, c = r.render_accumulate()
    assert beq(s, want_s) and (c == want_c).all()
    assert r.stats()["kernel_launches"] > one_batch, "the injected failure must have forced more than one batch"


def test_two_gpu_distributed_render_matches_one_gpu(rt):
    """scripts/dist_check.py under torchrun on 2 GPUs (NCCL): tile sharding bit-identical to the 1-GPU frame, sample
    sharding within 1e-5, with and without a sharded + all-gathered photon map.  Skipped on a single-GPU box."""
    import subprocess, sys
    from conftest import ROOT
    if rt.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          f"{ROOT}/scripts/dist_check.py"], capture_output=True, text=True, timeout=600)
    assert "DIST_CHECK PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_python_cli_writes_the_same_file_as_the_cxx_cli(rt, tmp_path):
    """`python -m ray_tracing_engine_b200.render_cli` (the multi-GPU front end, here on one GPU) and `bin/RayTracer`
    write byte-identical PPMs for the same command line."""
    import os, subprocess, sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from bench import materialize_meshes
    exe = os.path.join(ROOT, "ray-tracing-engine_b200", "bin", "RayTracer")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    mdir = materialize_meshes(str(tmp_path / "meshes"))
    args = ["-width", "96", "-height", "64", "-m", "1", "-N", "3", "-p", "2000", "-k", "4", "-meshdir", mdir, "-seed", "5"]
    a = subprocess.run([exe] + args + ["-o", "a.ppm"], capture_output=True, text=True, cwd=tmp_path)
    assert a.returncode == 0, a.stderr[-1000:]
    env = dict(os.environ, PYTHONPATH=ROOT)
    b = subprocess.run([sys.executable, "-m", "ray_tracing_engine_b200.render_cli"] + args + ["-o", "b.ppm"],
                       capture_output=True, text=True, cwd=tmp_path, env=env)
    assert b.returncode == 0, b.stderr[-1000:]
    assert open(tmp_path / "a.ppm", "rb").read() == open(tmp_path / "b.ppm", "rb").read()


def test_reference_program_with_the_binding_writes_the_same_file_as_our_cli(rt, tmp_path):
    """The drop-in, proven: oracle/_ref/RayTracer_b200_binding is the REFERENCE'S OWN program (its command line, OFF
    loader, scene assembly, background, savePPM) in which the one call `renderer.render(image)` (Main.cpp:224) goes
    through rt_render via the binding of INTEGRATION.md section 2 (oracle/ref_binding.h, built by oracle/Makefile).  Its
    output.ppm must be byte-identical to what bin/RayTracer -- the host written from scratch -- produces from the same
    .off files and arguments, without and with a photon map."""
    import os, subprocess
    from conftest import ROOT
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    binding, exe = os.path.join(ref_dir, "RayTracer_b200_binding"), os.path.join(ROOT, "ray-tracing-engine_b200", "bin", "RayTracer")
    if not (os.path.exists(binding) and os.path.exists(exe) and os.path.exists(os.path.join(ref_dir, "meshes", "cube_tri.off"))):
        pytest.skip("the binding is built where /root/reference exists and travels with the snapshot; not here")
    os.makedirs(os.path.join(ref_dir, "build"), exist_ok=True)
    for extra in ([], ["-p", "2000", "-k", "4"]):
        args = ["-width", "96", "-height", "64", "-m", "1", "-N", "3"] + extra
        a = subprocess.run([binding] + args + ["-o", str(tmp_path / "binding.ppm")], capture_output=True, text=True,
                           cwd=os.path.join(ref_dir, "build"))  # the reference resolves ../meshes/ from its cwd
        assert a.returncode == 0, a.stdout[-500:] + a.stderr[-1000:]
        b = subprocess.run([exe] + args + ["-meshdir", os.path.join(ref_dir, "meshes"), "-o", str(tmp_path / "ours.ppm")],
                           capture_output=True, text=True, cwd=tmp_path)
        assert b.returncode == 0, b.stderr[-1000:]
        x, y = open(tmp_path / "binding.ppm", "rb").read(), open(tmp_path / "ours.ppm", "rb").read()
        assert len(x) > 96 * 64 * 3 * 2 and x == y, "the reference program + binding and our CLI wrote different files"


def test_non_finite_vertices_are_rejected(rt):
    """A NaN or infinite coordinate would poison every box of the BVH; rt_create refuses the scene."""
    stock = rt.Scene.load(scene_path("stock"))
    for bad in (np.nan, np.inf, -np.inf, 3e8):
        pos = stock.pos.copy()
        pos[5, 1] = bad
        scene = rt.Scene(pos, stock.nrm, stock.tri, stock.mesh_tri_off, stock.mesh_vtx_off, stock.mats, stock.lights,
                         stock.cam, stock.w, stock.h, stock.lights_ctor)
        with pytest.raises(rt.RtError) as e:
            rt.Renderer(scene, 1, 0)
        assert e.value.code == -1 and "finite" in str(e.value)
