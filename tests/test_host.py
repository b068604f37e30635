"""CPU tests of the boundary and the host side (no GPU, no compute calls)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLD, ROOT, scene_path

import ray_tracing_engine_b200 as rt
from ray_tracing_engine_b200 import _capi

MESH_DIRS = ["/root/reference/meshes", os.path.join(ROOT, "oracle", "_ref", "meshes")]
MESHES = next((d for d in MESH_DIRS if os.path.exists(os.path.join(d, "cube_tri.off"))), None)
needs_meshes = pytest.mark.skipif(MESHES is None, reason="the reference's .off assets are not available here")


def beq(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and bool((a.view(np.uint32) == b.view(np.uint32)).all())


# ----------------------------------------------------------------------------- the C ABI
def header_functions():
    src = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    declared = header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/rt_b200.h but not exported"
    assert sorted(_capi.SYMBOLS) == declared, "the ctypes table and the header disagree"


def test_struct_layouts_match_the_header():
    assert C.sizeof(_capi.rt_material) == 32 and C.sizeof(_capi.rt_light) == 84 and C.sizeof(_capi.rt_camera) == 48
    assert C.sizeof(_capi.rt_params) == 64 and _capi.rt_params.seed.offset == 24
    assert C.sizeof(_capi.rt_scene) == 16 + 7 * 8 + 48
    assert _capi.rt_stats.device_ms.offset == 64 and C.sizeof(_capi.rt_stats) == 136 + 2 * 8 * len(_capi.KERNEL_CLASSES) + 8
    # ... and the library's own sizeof of every struct (rt_abi_sizes) agrees with the ctypes mirrors
    sizes = np.zeros(9, np.int32)
    assert _capi.load().rt_abi_sizes(_capi.ptr(sizes), 9) == 0
    ray, hit = np.dtype([("o", np.float32, 3), ("d", np.float32, 3)]), np.dtype([("t", np.int32), ("uvt", np.float32, 3)])
    mirrors = [C.sizeof(_capi.rt_material), C.sizeof(_capi.rt_light), C.sizeof(_capi.rt_camera), C.sizeof(_capi.rt_scene),
               C.sizeof(_capi.rt_params), ray.itemsize, hit.itemsize, 28, C.sizeof(_capi.rt_stats)]
    assert sizes.tolist() == mirrors, (sizes.tolist(), mirrors)


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (it never routes through the oracle)."""
    if rt.device_count() > 0:
        pytest.skip("a GPU is present")
    scene = rt.Scene.load(scene_path("stock"))
    with pytest.raises(rt.RtError) as e:
        rt.Renderer(scene, 1, 0)
    assert e.value.code == _capi.RT_ERR_NO_DEVICE
    src = "".join(open(os.path.join(ROOT, "ray-tracing-engine_b200", f)).read()
                  for f in ("__init__.py", "_capi.py", "distributed.py", "render_cli.py"))
    assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# ", ""), \
        "the product package must not import the oracle"
    out = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_parameter_validation():
    lib = _capi.load()
    scene = rt.Scene.load(scene_path("stock"))
    cs = scene._as_c()
    ctx = C.c_void_p()
    for kw, text in ((dict(width=0), "width"), (dict(num_photons=10, k=0), "k must"), (dict(num_photons=10, k=_capi.RT_MAX_K + 1), "k must"),
                     (dict(shard_count=2, shard_rank=2), "shard_rank"), (dict(num_rays=4, sample_first=3, sample_count=2), "sample")):
        base = dict(width=8, height=8, num_rays=1, mode=0)
        base.update(kw)
        p = rt._params(**base)
        assert lib.rt_create(C.byref(cs), C.byref(p), 0, C.byref(ctx)) == _capi.RT_ERR_INVALID
        assert text in lib.rt_last_error().decode()
    assert lib.rt_render(None, None) == _capi.RT_ERR_INVALID


def test_composite_formula():
    """rt_composite is host code: saveImage = update/N + background*(N-counter)/N (Renderer.cpp:262-265)."""
    from oracle import oracle as O
    g = np.random.default_rng(0)
    s = g.uniform(0, 5, (7, 9, 3)).astype(np.float32)
    c = g.integers(0, 6, (7, 9)).astype(np.int32)
    bg = g.uniform(0, 1, (7, 9, 3)).astype(np.float32)
    assert beq(rt.Renderer.composite(5, s, c, bg), O.PortOracle().composite(5, s, c, bg))


def test_shard_pixels_partition():
    for (w, h, n, tile) in ((420, 420, 8, 16), (100, 70, 3, 16), (33, 17, 4, 8), (5, 5, 7, 16)):
        parts = [rt.shard_pixels(w, h, r, n, tile) for r in range(n)]
        allpix = np.concatenate(parts)
        assert len(allpix) == w * h and len(np.unique(allpix)) == w * h
        for r, part in enumerate(parts):
            x, y = part % w, part // w
            assert ((((y // tile) * ((w + tile - 1) // tile) + x // tile) % n) == r).all()
    assert len(rt.shard_pixels(64, 64, 0, 1)) == 64 * 64


# ----------------------------------------------------------------------------- host mirror
def test_scene_file_roundtrip(tmp_path):
    s = rt.Scene.load(scene_path("lowres"))
    p = str(tmp_path / "x.rtscene")
    s.save(p)
    t = rt.Scene.load(p)
    for k in ("pos", "nrm", "mats", "lights", "cam"):
        assert beq(getattr(s, k), getattr(t, k))
    assert (s.tri == t.tri).all() and (s.tri_mesh() == t.tri_mesh()).all() and s.T == 1222


@needs_meshes
@pytest.mark.parametrize("name,off", [("stock", None), ("lowres", "example_low_res.off"), ("example", "example.off")])
def test_host_scene_assembly_matches_reference(name, off):
    """host/scene_host.cpp (OFF loader, normals, rotation, Cornell box, lights, camera) vs the scene the
    reference's own main() code assembles (golden dump): every float bit-identical."""
    g = rt.Scene.load(scene_path(name))
    s = rt.Scene.build(420, 420, MESHES, os.path.join(MESHES, off) if off else None)
    for k in ("pos", "nrm", "mats", "lights", "cam"):
        assert beq(getattr(s, k), getattr(g, k)), k
    assert (s.tri == g.tri).all() and (s.mesh_tri_off == g.mesh_tri_off).all() and (s.mesh_vtx_off == g.mesh_vtx_off).all()


def test_camera_for_other_aspect(gold):
    cam = rt.Scene.load(scene_path("stock")).with_size(380, 270).cam
    assert beq(cam, gold("camera_380x270.npz")["cam"])
    # golden values of SURVEY.md 8a-C for aspect 1
    np.testing.assert_allclose(rt.Scene.load(scene_path("stock")).cam[3:9],
                               [-0.379017353, -0.209387183, 1.55804682, 1.14500153, 0, -0.149348021], rtol=2e-7)


def test_off_loader_on_own_fixture(gold, tmp_path):
    g = gold("fixture_mixed_loaded.npz")
    pos, nrm, tri = rt.load_off(os.path.join(GOLD, "fixture_mixed.off"))
    assert beq(pos, g["pos"]) and beq(nrm, g["nrm"]) and (tri == g["tri"]).all()
    assert len(tri) == 1 + 2 + 3 + 1  # triangle, quad -> 2, pentagon -> 3, triangle
    with pytest.raises(RuntimeError) as e:
        rt.load_off(str(tmp_path / "missing.off"))
    assert "Error Loading OFF file: Error loading OFF file:" in str(e.value)  # Mesh.h:62-64,84-88
    # CRLF line endings, as in the reference's example meshes (which carry no comment line: the
    # reference's skipHashCommentLine does not step over '\r', Mesh.h:126-134, and neither do we)
    lines = [l for l in open(os.path.join(GOLD, "fixture_mixed.off"), "rb").read().split(b"\n") if not l.startswith(b"#")]
    crlf = tmp_path / "crlf.off"
    crlf.write_bytes(b"\r\n".join(lines))
    p2, n2, t2 = rt.load_off(str(crlf))
    assert beq(p2, pos) and (t2 == tri).all()


def test_subdivision_counts_and_midpoints():
    pos0, _, tri0 = rt.load_off(os.path.join(GOLD, "fixture_mixed.off"))
    pos, nrm, tri = rt.load_off(os.path.join(GOLD, "fixture_mixed.off"), subdivisions=2)
    assert len(tri) == 16 * len(tri0)
    assert beq(pos[:len(pos0)], pos0)
    a, b = pos0[tri0[0][0]], pos0[tri0[0][1]]
    assert beq(pos[len(pos0)], np.float32(0.5) * (a + b))  # first new vertex = midpoint of the first edge
    np.testing.assert_allclose(np.linalg.norm(nrm, axis=1), 1, atol=1e-6)


@needs_meshes
def test_synthetic_million_triangle_scene():
    s = rt.Scene.build(1920, 1080, MESHES, os.path.join(MESHES, "example_low_res.off"), 5)
    assert s.T == 1200 * 4 ** 5 + 22 and s.V == 614785 + 28  # SURVEY.md 8d cfg 5


def test_image_background_and_ppm(gold, tmp_path):
    g = gold("background.npz")
    assert beq(rt.Image(420, 420).fillBackground().pixels[:, 0, :], g["rows_420"])
    assert beq(rt.Image(380, 270).fillBackground().pixels[:, 7, :], g["rows_270"])
    img = rt.Image(3, 2)
    img.pixels[:] = np.float32([[[0, 0.5, 1], [0.999, 0.2, 0.1], [1, 1, 1]], [[0.004, 0.0039, 0.3], [0.25, 0.75, 0.6], [0, 0, 0]]])
    path = str(tmp_path / "a.ppm")
    img.savePPM(path)
    assert open(path).read() == "P3\n3 2\n255\n0 127 255 254 51 25 255 255 255 1 0 76 63 191 153 0 0 0 \n"
    # the C++ host writer produces the same bytes
    lib = rt._host_lib()
    p2 = str(tmp_path / "b.ppm")
    lib.rth_save_ppm(p2.encode(), 3, 2, _capi.ptr(img.pixels))
    assert open(p2).read() == open(path).read()
    # -p6 1: binary P6 with the same quantisation
    p3 = str(tmp_path / "c.ppm")
    lib.rth_save_ppm_binary(p3.encode(), 3, 2, _capi.ptr(img.pixels))
    assert open(p3, "rb").read() == b"P6\n3 2\n255\n" + bytes([0, 127, 255, 254, 51, 25, 255, 255, 255, 1, 0, 76, 63, 191, 153,
                                                                0, 0, 0])


def test_cli_banner_and_errors(tmp_path):
    """bin/RayTracer keeps the reference's command line (CommandLine.h:47-97): banner, defaults, errors."""
    exe = os.path.join(ROOT, "ray-tracing-engine_b200", "bin", "RayTracer")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    r = subprocess.run([exe, "-bogus", "1"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "Unknown argument <-bogus>" in r.stderr and "USAGE:" in r.stderr
    r = subprocess.run([exe, "-width"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "Missing argument" in r.stderr
    r = subprocess.run([exe, "-help"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and "USAGE:" in r.stderr
    r = subprocess.run([exe, "-m", "7", "-p", "100", "-k", "3", "-meshdir", "/nonexistent"], capture_output=True,
                       text=True, cwd=tmp_path)
    assert "Mode: Ray tracing" in r.stdout and "Photon map ON with 100 photons. Number of searched neighbours equals 3" in r.stdout
    assert "width: 380, height: 270" in r.stdout and "Output image filename: output.ppm" in r.stdout
    assert r.returncode == 1 and "Error Loading OFF file: Error loading OFF file: /nonexistent/cube_tri.off" in r.stderr
    # the additive options parse (and leave the reference's banner alone); without a GPU the run then stops in rt_create
    r = subprocess.run([exe, "-update", "4", "-p6", "1", "-knn", "exact", "-seed", "7", "-subdiv", "0", "-device", "0",
                        "-brute", "0", "-meshdir", "/nonexistent"], capture_output=True, text=True, cwd=tmp_path)
    assert "Mode: Ray tracing" in r.stdout and "Photon map OFF" in r.stdout and "Unknown argument" not in r.stderr
    assert r.returncode == 1 and "cube_tri.off" in r.stderr


def test_multi_gpu_cli_parses_the_reference_flags():
    from ray_tracing_engine_b200 import render_cli
    a = render_cli.parse(["-w", "420", "-height", "300", "-n", "8", "-m", "7", "-p", "100", "-k", "3", "-o", "x.ppm"])
    assert (a["width"], a["height"], a["numRays"], a["mode"], a["numPhotons"], a["k"], a["output"]) == \
        (420, 300, 8, 0, 100, 3, "x.ppm")           # an invalid mode falls back to ray tracing (CommandLine.h:84-87)
    d = render_cli.parse([])
    assert (d["width"], d["height"], d["numRays"], d["mode"], d["k"], d["output"]) == (380, 270, 16, 0, 5, "output.ppm")
    for bad, msg in ((["-bogus", "1"], "Unknown argument <-bogus>"), (["-width"], "Missing argument")):
        with pytest.raises(SystemExit) as e:
            render_cli.parse(bad)
        assert msg in str(e.value)


# ----------------------------------------------------------------------------- host BVH builder
def _check_bvh_policy(scene, nodes, slots):
    """Every internal node of every per-mesh tree obeys BVH::from_triangles (source/BVH.h:100-161):
    leaf iff one triangle (:123); cut axis = first axis of strictly largest extent of the node's vertex box
    (:131-140); left child = the floor(n/2) triangles with the smallest key sum_v vertex[v][axis] (:141-158;
    ties may fall on either side, the reference's std::sort is unstable).  Child boxes contain their vertices."""
    P, T = scene.pos.reshape(-1, 3), scene.tri.reshape(-1, 3)
    assert sorted(slots.tolist()) == list(range(scene.T)), "every triangle sits in exactly one leaf slot"
    tri_mesh = scene.tri_mesh()
    refs = nodes[:, 12:14].copy().view(np.int32)
    checked = [0]

    def walk(ref):  # -> array of global triangle ids below ref
        if ref < 0:
            return np.array([slots[~ref]])
        l, r = walk(refs[ref, 0]), walk(refs[ref, 1])
        for side, tris in ((0, l), (1, r)):
            v = P[T[tris].reshape(-1)]
            lo, hi = nodes[ref, 6 * side:6 * side + 3], nodes[ref, 6 * side + 3:6 * side + 6]
            assert (lo <= v.min(0)).all() and (hi >= v.max(0)).all(), "child box does not contain its triangles"
        both = np.concatenate([l, r])
        if len(set(tri_mesh[both].tolist())) == 1:  # a node of a per-mesh tree (not the top-level join)
            v = P[T[both].reshape(-1)]
            ext = v.max(0) - v.min(0)
            axis, longest = 0, np.float32(0)
            for a in range(3):
                if ext[a] > longest:
                    axis, longest = a, ext[a]
            key = lambda ids: (P[T[ids, 0], axis] + P[T[ids, 1], axis]) + P[T[ids, 2], axis]
            assert len(l) == len(both) // 2, "left child must hold floor(n/2) triangles"
            assert key(l).max() <= key(r).min(), f"node {ref}: not a median split on axis {axis}"
            checked[0] += 1
        return both

    assert len(walk(0)) == scene.T
    return checked[0]


@pytest.mark.parametrize("name", ["stock", "lowres", "example"])
def test_host_bvh_follows_the_reference_split_policy(name):
    path = scene_path(name)
    if not os.path.exists(path):
        pytest.skip(f"{name} scene fixture not present")
    scene = rt.Scene.load(path)
    nodes, slots, depth = scene.build_bvh_host()
    nonempty = int((np.diff(scene.mesh_tri_off) > 0).sum())
    assert nodes.shape[0] == scene.T - 1, "a full binary tree with one-triangle leaves has T-1 internal nodes"
    n_checked = _check_bvh_policy(scene, nodes, slots)
    assert n_checked == scene.T - nonempty
    biggest = int(np.diff(scene.mesh_tri_off).max())
    assert depth >= int(np.ceil(np.log2(biggest))) and depth <= int(np.ceil(np.log2(biggest))) + 5


def test_restated_libm_sinf_cosf_equals_this_machines_libm(tmp_path):
    """RayTracer.h:104-106 calls libm's binary32 cos/sin, which are not correctly rounded; the device restates glibc's
    algorithm (csrc/rt_device.cuh: libm_sincosf).  The same restatement in C must reproduce sinf/cosf of the C library
    here bit for bit on every 5th binary32 in [2^-15, 2 pi] (the arguments hsphereUniformSample can produce; the
    exhaustive run over all 147 M values was done once: 0 mismatches)."""
    exe = str(tmp_path / "libm_check")
    src = os.path.join(ROOT, "tests", "libm_sincosf_check.c")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-o", exe, src, "-lm"], check=True)
    out = subprocess.run([exe, "38000000", "40c90fdc", "5"], capture_output=True, text=True, check=True).stdout.split()
    n, bad_sin, bad_cos = (int(v) for v in out)
    assert n > 29_000_000 and bad_sin == 0 and bad_cos == 0, out


def test_pointcloud_writer_matches_the_references(tmp_path):
    """PhotonMap::saveToPCD (PhotonMap.h:59-84): the C++ host writer (the CLI's own, also behind
    Renderer.savePhotonMap) produces byte for byte the file the reference's writer produced for the same particles
    (tests/golden/pointcloud_golden.pcd, generated by oracle/gen_golden.py: pointcloud) -- and, where the reference
    build is present, what it produces right now."""
    plist = np.load(os.path.join(GOLD, "pointcloud_input.npy"))
    out = str(tmp_path / "ours.pcd")
    rt.save_pcd(out, plist)
    want = open(os.path.join(GOLD, "pointcloud_golden.pcd"), "rb").read()
    assert open(out, "rb").read() == want
    from oracle import oracle as O
    if O.have_ref():
        live = str(tmp_path / "ref.pcd")
        O.RefOracle().save_pcd(plist, live)
        assert open(live, "rb").read() == want
    rt.save_pcd(out, np.zeros((0, 7), np.float32))  # the reference's own call writes this header-only file (164 bytes)
    assert len(open(out, "rb").read()) == 164


@needs_meshes
def test_input_list_and_binary_off_cache(tmp_path):
    """-i a.off,b.off,c.off: the first two replace mesh_cube / mesh_cube2, further ones are appended in order; the
    binary OFF cache returns bit-identical meshes, is keyed by file size + mtime + subdivisions, and survives a
    corrupt cache file."""
    low, cube, cube2 = (os.path.join(MESHES, f) for f in ("example_low_res.off", "cube_tri.off", "cube_tri2.off"))
    stock = rt.Scene.build(64, 64, MESHES)
    one = rt.Scene.build(64, 64, MESHES, low)
    assert one.M == 5 and one.T == stock.T - 12 + 1200
    same = rt.Scene.build(64, 64, MESHES, f"{cube},{cube2}")
    assert beq(same.pos, stock.pos) and (same.tri == stock.tri).all() and beq(same.nrm, stock.nrm)
    three = rt.Scene.build(64, 64, MESHES, f"{low},{cube2},{cube}")
    assert three.M == 6 and three.T == one.T + 12
    assert beq(three.pos[:one.V], one.pos) and (three.tri[:one.T] == one.tri).all()
    # the appended mesh gets mesh_cube's material and rotation: it is the stock scene's mesh 3
    a, b = stock.mesh_vtx_off[3], stock.mesh_vtx_off[4]
    assert beq(three.pos[one.V:], stock.pos[a:b]) and beq(three.mats[5], stock.mats[3])
    # cache: first build writes, second reads; both equal the uncached build
    cache = str(tmp_path / "cache")
    plain = rt.Scene.build(64, 64, MESHES, low, 2)
    first = rt.Scene.build(64, 64, MESHES, low, 2, cache)
    files = sorted(os.listdir(cache))
    assert len(files) == 2 and all(f.endswith(".offbin") for f in files) and any(".s2." in f for f in files)
    second = rt.Scene.build(64, 64, MESHES, low, 2, cache)
    for s in (first, second):
        assert beq(s.pos, plain.pos) and beq(s.nrm, plain.nrm) and (s.tri == plain.tri).all()
    with open(os.path.join(cache, [f for f in files if ".s2." in f][0]), "r+b") as fh:  # truncate: must be re-parsed
        fh.truncate(100)
    third = rt.Scene.build(64, 64, MESHES, low, 2, cache)
    assert beq(third.pos, plain.pos) and (third.tri == plain.tri).all()
    with pytest.raises(RuntimeError):
        rt.Scene.build(64, 64, MESHES, "/nonexistent/x.off", 0, cache)


def test_host_kdtree_builders(gold):
    """rt_build_kdtree_host: (1) the default builder reproduces the reference's own tree (kdtree::make_tree with
    libstdc++'s nth_element) on the golden photon list; (2) the canonical builder of the exact k-NN mode is a valid
    median-split kd-tree, orders equal coordinates by list index, and does not depend on the input order of ties."""
    g = gold("photons.npz")
    nodes, _, h = rt.build_kdtree_host(g["list"])
    assert beq(nodes, g["nodes"]) and h == int(np.floor(np.log2(len(nodes)))) + 1
    lst = g["list"].copy()
    lst[100:400, 0] = lst[100, 0]  # many equal x coordinates
    can, orig, h2 = rt.build_kdtree_host(lst, canonical=True)
    assert h2 == h and sorted(orig.tolist()) == list(range(len(lst))) and beq(can, lst[orig])

    def check(b, e, axis):
        if e <= b:
            return
        n = b + (e - b) // 2
        key = lambda i: (can[i, axis], orig[i])
        assert all(key(i) < key(n) for i in range(b, n)) and all(key(n) < key(i) for i in range(n + 1, e))
        check(b, n, (axis + 1) % 3)
        check(n + 1, e, (axis + 1) % 3)
    check(0, len(can), 0)
