"""Where does the per-rank time of a ring-sharded spectrum go?  torchrun --nproc-per-node N scripts/strong_probe.py
Variants of the bench's strong-scaling step, phases per rank (CUDA events on the library's stream) and wall time."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from radlite_b200 import shard, synth  # noqa: E402
from radlite_b200.api import Renderer  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
m = synth.config(2)
nl, nfr = m.nlines, m.nfr
g = Renderer(local)
g.load_model(m)
nrr, nphi, nray = g.camera_dims()
cost = torch.zeros(nrr + 1, dtype=torch.float64, device=dev)
if rank == 0:
    cost = torch.from_numpy(g.plan_costs(1, nl, nfr, m.passband)).to(dev)
dist.broadcast(cost, src=0)
blocks = shard.split_rings(nrr, world, cost.cpu().numpy())
lo, hi = blocks[rank]
rs = torch.zeros((nl, nrr + 1, nfr), dtype=torch.float64, device=dev)


def run(name, reduce, sync_each, K=6):
    def step():
        g.invalidate_geometry()
        t = g.render_rings_device(1, nl, nfr, m.passband, synth.PARSEC, lo, hi, rs.data_ptr())
        if reduce:
            dist.reduce(rs, dst=0, op=dist.ReduceOp.SUM)
        if sync_each:
            dist.barrier()
            torch.cuda.synchronize()
        return t
    for _ in range(3):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ms = np.zeros(5)
    for _ in range(K):
        ms += np.array(step())
    dist.barrier()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / K * 1e3
    ms /= K
    out = [None] * world
    dist.all_gather_object(out, (rank, lo, hi, [round(float(x), 3) for x in ms], round(wall, 3)))
    if rank == 0:
        print(f"== {name}", flush=True)
        for o in out:
            print("   rank %d rings %d..%d geometry/prep/integrate/flux/total %s wall %.3f" % o, flush=True)


run("A reduce, free running", True, False)
run("B no exchange, free running", False, False)
run("C reduce, barrier after each step", True, True)
run("D no exchange, barrier after each step", False, True)
dist.barrier()
dist.destroy_process_group()
