"""World-size-2 tests of the multi-GPU host logic on CPU (gloo).  The renderer itself needs a GPU, so the
per-rank work is produced by the CPU oracle (tests may use it as a stand-in); what is under test is the
product's sharding / all-gather / reduce plumbing in ray-tracing-engine_b200/distributed.py."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, scene_path


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    import ray_tracing_engine_b200 as rt
    from ray_tracing_engine_b200 import distributed as D

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = O.FlatScene.load(scene_path("stock"))
    W, H, N = 40, 24, 4
    flat.w, flat.h = W, H
    port_o = O.PortOracle(flat)

    # photons: sharded emission, all-gather, identical list on every rank
    per = 900 // 3
    first, count = D.path_range(per, rank, world)
    shard = port_o.photon_map_create(900, 5, first, count)
    local, counts = shard.get()[0], shard.light_counts(3)
    assert counts.sum() == len(local)
    full_want = port_o.photon_map_create(900, 5).get()[0]
    gathered = D.gather_photons(local, counts)
    assert gathered.shape == full_want.shape and (gathered.view(np.uint32) == full_want.view(np.uint32)).all(), \
        "the spliced shards must equal the list the single-process run builds, in (light, path) order"

    # frame: tile shards (exact) and sample shards, reduced to rank 0
    want = port_o.render(N, 1, 3)
    for shard in ("tile", "sample"):
        if shard == "tile":
            mine = rt.shard_pixels(W, H, rank, world, 8)
            part = port_o.render(N, 1, 3)
            mask = np.zeros(W * H, bool)
            mask[mine] = True
            s = np.where(mask.reshape(H, W, 1), part["sum_rgb"], 0).astype(np.float32)
            c = np.where(mask.reshape(H, W), part["counter"], 0).astype(np.int32)
        else:
            f, n = D.sample_range(N, rank, world)
            part = port_o.render(N, 1, 3, samples=(f, f + n))
            s, c = part["sum_rgb"], part["counter"]
        st, ct = torch.from_numpy(s.copy()), torch.from_numpy(c.copy())
        D.reduce_frame(st, ct, 0)
        if rank == 0:
            assert (ct.numpy() == want["counter"]).all()
            if shard == "tile":
                assert (st.numpy().view(np.uint32) == want["sum_rgb"].view(np.uint32)).all(), "tile sharding is exact"
            else:
                np.testing.assert_allclose(st.numpy(), want["sum_rgb"], rtol=1e-6, atol=1e-6)
    dist.barrier()
    dist.destroy_process_group()
    out.put((rank, "ok"))


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5)[0] for _ in range(2)) == [0, 1]


def test_ranges_cover_everything():
    sys.path.insert(0, ROOT)
    from ray_tracing_engine_b200 import distributed as D
    for total, world in ((16666, 8), (128, 3), (5, 8), (0, 2)):
        rs = [D.path_range(total, r, world) for r in range(world)]
        assert rs[0][0] == 0 and sum(c for _, c in rs) == total
        assert all(rs[i][0] + rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        assert [D.sample_range(total, r, world) for r in range(world)] == rs


def test_empty_sample_shards_render_nothing():
    """More ranks than samples: rt_params reads sample_count <= 0 as "all samples", so an empty share must be
    expressed as the empty range at num_rays (found by the 8-GPU run of scripts/dist_check.py with N = 2)."""
    sys.path.insert(0, ROOT)
    from ray_tracing_engine_b200 import distributed as D
    for total, world in ((2, 8), (5, 8), (128, 8), (1, 2)):
        covered = []
        for r in range(world):
            kw = D.sample_shard_kwargs(total, r, world)
            if kw["sample_count"] == 0:
                assert kw["sample_first"] == total, "an empty share must not collapse to 'all samples'"
            else:
                covered += list(range(kw["sample_first"], kw["sample_first"] + kw["sample_count"]))
        assert covered == list(range(total))
