"""scripts/dist_check.py on ONE GPU: the ranks of a `world`-rank distributed render run one after the other in this
process (same contexts, same device-resident photon path, same packed frames; the NCCL collectives are replaced by
torch.cat / a sum), against the single-context frame.  Catches everything that is not NCCL itself."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import ray_tracing_engine_b200 as rt
from ray_tracing_engine_b200 import distributed as D

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
scene = rt.Scene.load(os.path.join(ROOT, "tests/golden/scenes/stock.rtscene"))
W = H = 200
scene.w = scene.h = W
bg = rt.Image(W, H).fillBackground().pixels
ok = True
for rep in range(int(os.environ.get("REPEAT", "2"))):
  for (N, mode, photons, k) in ((8, 1, 0, 5), (2, 0, 30000, 10), (4, 1, 30000, 7)):
    for shard in ("tile", "sample"):
        rs = []
        for q in range(world):
            kw = dict(seed=5, device=0)
            kw.update(dict(shard_rank=q, shard_count=world) if shard == "tile" else D.sample_shard_kwargs(N, q, world))
            rs.append(rt.Renderer(scene, N, mode, None, photons, k, **kw))
        if photons:
            per, L = rs[0].photons_per_light(), scene.L
            cap = max(1, max(D.path_range(per, q, world)[1] for q in range(world)) * L)
            gathered = torch.empty((world * cap, 7), dtype=torch.float32, device=dev)
            counts = np.zeros((world, L), np.int64)
            for q in range(world):
                local = torch.empty((cap, 7), dtype=torch.float32, device=dev)
                first, count = D.path_range(per, q, world)
                counts[q], _ = rs[q].emit_photons_device(first, count, local.data_ptr(), cap)
                gathered[q * cap:(q + 1) * cap] = local
            torch.cuda.synchronize()
            total = int(counts.sum())
            for q in range(world):
                out = torch.empty((max(total, 1), 7), dtype=torch.float32, device=dev)
                assert rs[q].splice_photons_device(gathered.data_ptr(), world, cap, counts, out.data_ptr(), max(total, 1)) == total
                rs[q].set_photons_device(out.data_ptr(), total)
        acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
        for q in range(world):
            packed = torch.empty((H, W, 4), dtype=torch.float32, device=dev)
            torch.cuda.synchronize()
            rs[q].render_accumulate_packed_device(packed.data_ptr())
            acc += packed
        torch.cuda.synchronize()
        img = rs[0].composite_packed_device(N, acc.data_ptr(), bg)
        for r in rs:
            r.close()
        r1 = rt.Renderer(scene, N, mode, None, photons, k, seed=5, device=0)
        want = r1.render(rt.Image(W, H).fillBackground()).pixels
        r1.close()
        same = (img.view(np.uint32) == want.view(np.uint32)).all(axis=-1)
        err = float(np.abs(img - want).max())
        good = bool(same.all()) if shard == "tile" else err < 1e-5
        ok &= good
        bad = np.argwhere(~same)
        print(f"rep {rep} N={N} mode={mode} photons={photons} shard={shard}: identical {same.mean():.6f} max|diff| {err:.2e} "
              f"{'OK' if good else 'FAIL'} finite img {np.isfinite(img).all()} want {np.isfinite(want).all()} "
              f"bad bbox {bad.min(0).tolist() if len(bad) else None}-{bad.max(0).tolist() if len(bad) else None} "
              f"img max {np.nanmax(img):.3g} want max {np.nanmax(want):.3g}", flush=True)
print("DIST_EMULATE", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
