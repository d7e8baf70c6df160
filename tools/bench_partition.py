#!/usr/bin/env python
"""Forward implicit steps of ONE sheet cut into strips over the GPUs of a node (include/tsl.h tsl_dist_init):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/bench_partition.py --rows R --cols M

The sheet has N*R vertex rows x (M+1) columns (weak scaling: R fixed per GPU; for strong scaling divide R by N).  Device-timed with CUDA
events, MAX over ranks; prints one JSON line on rank 0.  Forward steps only (the adjoint is not partitioned): this is NOT bench.py's
metric, it documents what the partitioned path costs next to `--gpus N` replicas."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thinshelllab_b200.synthetic import strip_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=708)
ap.add_argument("--cols", type=int, default=707)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
s = strip_scene(args.rows, args.cols, rank, world, device=dev)
e = s.engine
for _ in range(args.warmup):
    s.time_step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
c0 = e.dist_stats()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
log = []
for _ in range(args.steps):
    st = s.time_step()
    log.append({"newton": st.newton_iters, "pcg": st.linear_iters, "contacts_rank0": st.n_contacts, "converged": bool(st.converged), "flags": st.flags})
ev1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
c1 = e.dist_stats()
if rank == 0:
    tris = s.spec["n_tris_global"]
    print(json.dumps({"what": "strip-partitioned forward steps (one sheet over N GPUs)", "n_gpus": world, "sheet_vertex_rows": world * args.rows, "cols": args.cols,
                      "n_tris": tris, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms.item() / args.steps,
                      "tri_steps_per_s_forward_only": tris * args.steps / (ms.item() * 1e-3), "per_step": log,
                      "halo_exchanges": c1["halo_exchanges"] - c0["halo_exchanges"], "allreduces": c1["allreduces"] - c0["allreduces"]}))
if world > 1:
    dist.destroy_process_group()
