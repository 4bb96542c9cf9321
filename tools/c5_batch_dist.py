#!/usr/bin/env python
"""BASELINE configs[4] sharded over the GPUs along the PATH axis ("shard paths across GPUs instead when nf is
small"): every rank runs its contiguous block of slant paths through wsm.measurement_vecFromSensor (observer epilogue on
the device), then one NCCL all-reduce combines the channel sums.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/c5_batch_dist.py --paths 64
"""
import argparse
import copy
import gc
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arts_b200 import _abi as abi  # noqa: E402
from arts_b200 import shard, synth, wsm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--paths", type=int, default=64)
ap.add_argument("--channels", type=int, default=20)
ap.add_argument("--grid", type=int, default=50)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
wsm.set_device(local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

base = synth.case_c5_single()
tg = (("T",), ("VMR", 0))
nq = len(tg)
cat = wsm.Catalog(base.cat)
per = base.nf // args.channels
channels = [[(int(j), (1.0 / per, 0.0, 0.0, 0.0)) for j in range(c * per, (c + 1) * per)] for c in range(args.channels)]
nx = args.grid * nq + 1
path_map = []
for ip in range(base.np_):
    pos = ip * (args.grid - 1) / (base.np_ - 1)
    i0 = min(int(pos), args.grid - 2)
    path_map.append([[(t * args.grid + i0, 1.0 - (pos - i0)), (t * args.grid + i0 + 1, pos - i0)] for t in range(nq)])
obs = abi.Observer(nx=nx, path_map=path_map, bkg_T=288.0, bkg_rows=[(nx - 1, 1.0)], unit="PlanckBT", channels=channels)
zen = np.linspace(180.0, 120.0, args.paths)
off, cnt = shard.path_ranges(args.paths, world)[rank]
sims = []
for z in zen[off:off + cnt]:
    atm = copy.deepcopy(base.atm)
    atm.los = np.tile([z, 0.0], (base.np_, 1))
    sims.append((atm, base.r / abs(np.cos(np.deg2rad(z))), obs))
gc.disable()
wsm.measurement_vecFromSensor(cat, base.f, sims[:2], jac_targets=tg, hse_derivative=1)  # warm-up


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


barrier()
t0 = time.perf_counter()
y, J = wsm.measurement_vecFromSensor(cat, base.f, sims, jac_targets=tg, hse_derivative=1)
yd, Jd = torch.from_numpy(y).cuda(), torch.from_numpy(J).cuda()
if world > 1:
    shard.reduce_measurement(yd, Jd)
barrier()
t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    evals = float(len(base.cat.f0)) * base.nf * base.np_ * args.paths
    print(json.dumps({"workload": f"C5 batch sharded by path: {args.paths} paths over {world} GPU(s), T + VMR targets, observer epilogue",
                      "n_gpus": world, "seconds": t.item(), "paths_per_s": args.paths / t.item(), "evals_per_s": evals / t.item(),
                      "y_mean_K": float(yd.mean().item() / args.paths), "exchange": "one all-reduce of [channels] + [channels, nx] doubles"}))
if world > 1:
    dist.destroy_process_group()
