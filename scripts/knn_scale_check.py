"""k-NN parity at scale: GPU kd_knearest_sorted against the CPU oracle (kdtree::knearest restated) on the photon
map the GPU emitted, queries = primary hit points of the frame."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, ray_tracing_engine_b200 as rt
from oracle import oracle as O
path = os.path.join(ROOT, "tests/golden/scenes/stock.rtscene")
scene = rt.Scene.load(path)
port = O.PortOracle(O.FlatScene.load(path))
g = np.random.default_rng(5)
for photons, k, nq in ((50000, 10, 40000), (500000, 50, 20000)):
    r = rt.Renderer(scene, 1, 0, None, photons, k, seed=1)
    plist = r.emit_photons()[0]                # (light, path) order, as PhotonMap's constructor stores them
    r.set_photons(plist)
    nodes, left, right, root = r.kdtree()
    pm = port.photon_map_from_list(plist)     # same list in, same nth_element sequence -> same tree
    on, ol, orr, oroot = pm.layout()
    assert (on.view(np.uint32) == nodes.view(np.uint32)).all() and (ol == left).all()
    q = nodes[g.integers(0, len(nodes), nq), :3] + g.normal(size=(nq, 3)).astype(np.float32) * np.float32(0.02)
    q[: nq // 4] = nodes[g.integers(0, len(nodes), nq // 4), :3]  # exact photon positions: distance 0, ties
    t = time.time(); idx = r.knearest(q, k); tg = time.time() - t
    t = time.time(); out, visited, oidx = pm.knn(q, k, want_index=True); tc = time.time() - t
    same_order = (idx == oidx).all(axis=1)
    same_set = (np.sort(idx, 1) == np.sort(oidx, 1)).all(axis=1)
    same_pos = (nodes[idx][:, :, :3].view(np.uint32) == out[:, :, :3].view(np.uint32)).all(axis=(1, 2))
    print(f"p={photons} stored={len(nodes)} k={k}: {nq} queries, identical order {same_order.mean():.6f}, identical set "
          f"{same_set.mean():.6f}, identical positions {same_pos.mean():.6f}; gpu {tg:.3f}s cpu {tc:.3f}s, cpu visits/query {visited.mean():.0f}")
    bad = np.where(~same_order)[0][:3]
    for b in bad:
        d1 = np.linalg.norm(nodes[idx[b], :3] - q[b], axis=1); d2 = np.linalg.norm(nodes[oidx[b], :3] - q[b], axis=1)
        print("  query", b, "gpu", idx[b][:12], d1[:6], "cpu", oidx[b][:12], d2[:6])
