#!/usr/bin/env python
"""Parity of the strip-partitioned forward step against the same sheet on one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_parity.py [R M steps]

Every rank steps its strip; rank 0 then steps the whole sheet in one context (world = 1) and compares the owned rows of every rank.
Prints one JSON line; exit code 1 when the positions differ by more than 3e-7 m or the contact counts differ."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thinshelllab_b200.synthetic import strip_scene  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 16
M = int(sys.argv[2]) if len(sys.argv) > 2 else 24
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
s = strip_scene(R, M, rank, world, device=dev)
e, sp = s.engine, s.spec


def gather_owned():
    own = e.pos[sp["own0"]:sp["own1"]].clone()
    own[:, 0] += sp["x_shift"]                                # back to the global frame
    parts = [torch.empty_like(own) for _ in range(world)] if world > 1 else [own]
    if world > 1:
        dist.all_gather(parts, own)                           # equal-sized strips
    return torch.cat(parts).cpu().numpy()


log, traj = [], []
for k in range(steps):
    st = s.time_step()
    log.append((st.newton_iters, st.linear_iters, st.n_contacts, bool(st.converged), st.energy, st.flags))
    traj.append(gather_owned())
glob = traj[-1]
ok = True
out = {"world": world, "R": R, "M": M, "steps": steps, "per_step_rank0": log, "comms_rank0": e.dist_stats()}
if rank == 0:
    ref = strip_scene(world * R, M, 0, 1, device=dev)
    rlog, errs = [], []
    for k in range(steps):
        st = ref.time_step()
        rlog.append((st.newton_iters, st.linear_iters, st.n_contacts, bool(st.converged), st.energy, st.flags))
        x = ref.engine.pos[:world * R * (M + 1)].cpu().numpy().copy()
        x[:, 0] += ref.spec["x_shift"]
        errs.append(float(np.abs(traj[k] - x).max()))
    err = errs[-1]
    out.update(per_step_single_gpu=rlog, max_abs_pos_err_m=err, pos_err_per_step=errs)
    ok = err < 3e-7 and all(a[3] and b[3] for a, b in zip(log, rlog))
    print(json.dumps(out))
if world > 1:
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.broadcast(flag, 0)
    ok = flag.item() == 0
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
