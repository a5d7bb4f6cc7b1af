#!/usr/bin/env python
"""BASELINE configs[3]: 16-qubit Eagle circuit of depth 40, 10^5 trajectories, ONE backend.execute_sharded call split over the
ranks of the process group (torchrun); rank 0 prints one JSON line.  Timed on the device, max over ranks."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from noisycircuits_b200 import circuits, noise_io
from noisycircuits_b200.backend import execute_sharded
from noisycircuits_b200.solvers import RemoteExecutor

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
sq, tq, meas, _ = noise_io.load_noise_tables(os.path.join(ROOT, "tests", "golden", "eagle16_noise.npz"))
n, T = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 100000
instr = circuits.eagle_layered(n, 40, circuits.coupled_pairs(tq, n), seed=42)
ex = RemoteExecutor(n, sq, tq, 1, specialize=True)
execute_sharded(n, instr, sq, tq, meas, None, 64 * world, executor=ex)            # compile / load / warm up
best = None
for _ in range(2):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = execute_sharded(n, instr, sq, tq, meas, None, T, executor=ex)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    best = float(dt) if best is None else min(best, float(dt))
if rank == 0:
    print(json.dumps({"config": "C4 16q Eagle depth 40, one execute_sharded call", "n_gpus": world, "trajectories": T, "seconds": best,
                      "trajectories_per_s": T / best, "norm": float(np.sum(out)), "specialized_launches": ex.last_stats["specialized_launches"]}), flush=True)
if world > 1:
    dist.destroy_process_group()
