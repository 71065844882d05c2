"""Multi-GPU check (run under torchrun, one rank per GPU): line-block and ring-block sharded renders
over NCCL must be bit-identical to the single-GPU render.  The library picks its integrate kernel by the
number of lines per batch (ztile_kernel from 8, tile_kernel below; they differ at the 1e-13 level), so the
check pins one kernel at a time (rl_set_kernel: z, then tile) for the bitwise comparison and then verifies the
library's own choice to 1e-10.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/mgpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radlite_b200 import api, shard, synth  # noqa: E402
from radlite_b200.api import Renderer  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for kern in ("z", "tile", ""):
  api.DEFAULT_KERNEL = kern or "auto"
  same = np.array_equal if kern else (lambda a, b: np.allclose(a, b, rtol=1e-10, atol=0.0))
  m = synth.config(2, nr=60, nth=24, nphi=24, nrext=-12, nlines=11)
  g = Renderer(local)
  g.load_model(m)
  flux = shard.render_spectrum_sharded(g, m.nlines, m.nfr, m.passband, synth.PARSEC, rank, world, dev)
  m1 = synth.config(1, nr=60, nth=24, nphi=24, nrext=-12)
  g1 = Renderer(local)
  g1.load_model(m1)
  nrr, nphi, _ = g1.camera_dims()
  cube = np.zeros((1, nrr + 1, nphi, m1.nfr))
  f1 = shard.render_line_ring_sharded(g1, 1, 1, m1.nfr, m1.passband, synth.PARSEC, rank, world, dev, image=cube)
  t = torch.from_numpy(cube).to(dev)
  dist.reduce(t, dst=0)
  if rank == 0:
      ref = g.render(1, m.nlines, m.nfr, m.passband, synth.PARSEC)["flux"]
      ref1 = g1.render(1, 1, m1.nfr, m1.passband, synth.PARSEC, want_image=True)
      ok = (same(flux, ref), same(f1, ref1["flux"]), same(t.cpu().numpy(), ref1["image"]))
      print(f"MGPU_CHECK kernel={kern or 'auto'} world={world} line-sharded identical={ok[0]} ring-sharded identical={ok[1]} cube slabs identical={ok[2]}")
      assert all(ok)
dist.barrier()
dist.destroy_process_group()
