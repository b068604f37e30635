"""Multi-GPU front end: the reference's command line (source/CommandLine.h:9-102), one process per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        -m ray_tracing_engine_b200.render_cli -width 1920 -height 1080 -m 1 -N 1024 -i ../meshes/example.off

Same flags and defaults as `bin/RayTracer` (-w/-width, -h/-height, -o/-output, -N/-n/-numRays, -m/-mode,
-p/-numPhotons, -k; additive: -i/-input a.off[,b.off...], -meshdir, -cache, -subdiv, -seed, -p6, -shard tile|sample).  The scene is assembled by
the C++ host code (lib/librt_host.so), every rank renders its share (interleaved 16x16 tiles --
bit-identical to one GPU -- or sample-index ranges), photons are emitted sharded and all-gathered, the frame is
sum-reduced to rank 0 over NCCL, composited on its GPU and written as output.ppm.  Without torchrun it renders on one GPU.
"""
from __future__ import annotations

import os
import sys
import time


def parse(argv):
    a = dict(width=380, height=270, numRays=16, mode=0, numPhotons=0, k=5, output="output.ppm", input=None,
             meshdir="../meshes", subdiv=0, seed=1, p6=0, shard="tile", cache=None)
    names = {"-w": "width", "-width": "width", "-h": "height", "-height": "height", "-o": "output", "-output": "output",
             "-N": "numRays", "-n": "numRays", "-numRays": "numRays", "-m": "mode", "-mode": "mode", "-p": "numPhotons",
             "-numPhotons": "numPhotons", "-k": "k", "-i": "input", "-input": "input", "-meshdir": "meshdir",
             "-subdiv": "subdiv", "-seed": "seed", "-p6": "p6", "-shard": "shard", "-cache": "cache"}
    i = 0
    while i < len(argv):
        flag = argv[i]
        if i == len(argv) - 1:  # CommandLine.h:50-57
            raise SystemExit("USAGE: see the module docstring" if flag == "-help" else "Missing argument")
        if flag not in names:
            raise SystemExit(f"Unknown argument <{flag}>")
        key, val = names[flag], argv[i + 1]
        a[key] = val if key in ("output", "input", "meshdir", "shard", "cache") else int(val)
        i += 2
    if a["mode"] != 1:
        a["mode"] = 0  # CommandLine.h:84-87
    if a["shard"] not in ("tile", "sample"):
        raise SystemExit("-shard must be tile or sample")
    return a


def main(argv=None):
    a = parse(sys.argv[1:] if argv is None else argv)
    import numpy as np
    import torch
    import torch.distributed as dist
    import ray_tracing_engine_b200 as rt
    from ray_tracing_engine_b200 import distributed as D

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    own_group = world > 1 and not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.time()
    scene = rt.Scene.build(a["width"], a["height"], a["meshdir"], a["input"], a["subdiv"], a["cache"])
    background = rt.Image(a["width"], a["height"]).fillBackground().pixels
    if world > 1:
        img = D.render_distributed(scene, a["numRays"], a["mode"], a["numPhotons"], a["k"], background=background,
                                   seed=a["seed"], shard=a["shard"], device=dev, local_device=local)
    else:
        r = rt.Renderer(scene, a["numRays"], a["mode"], None, a["numPhotons"], a["k"], seed=a["seed"], device=local)
        out = rt.Image(a["width"], a["height"])
        out.pixels = background.copy()
        img = r.render(out).pixels
    if rank == 0:
        lib = rt._host_lib()
        buf = np.ascontiguousarray(img, np.float32)
        (lib.rth_save_ppm_binary if a["p6"] else lib.rth_save_ppm)(a["output"].encode(), a["width"], a["height"],
                                                                   rt._capi.ptr(buf))
        print(f"Total time is {int(time.time() - t0)}[s]  ({world} GPU{'s' if world > 1 else ''}, {a['shard']}-sharded)")
    if own_group:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
