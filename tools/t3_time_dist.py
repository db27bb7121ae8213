"""Per-round device timing of the sharded T3-k run (launch with torch.distributed.run)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import ceno_b200 as cb
from ceno_b200 import synth
from ceno_b200.dist import eq_slice_scalar

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = world.bit_length() - 1
kl, nl = k - g, 1 << (k - g)
dev = cb.Device(local)
w = synth.fill_ext(0xE9, k)
A = cb.MultilinearExtension.from_evaluations_ext_vec(dev, kl, synth.fill_ext(0xC0FFEE ^ 1, nl, start=rank * nl))
B = cb.MultilinearExtension.from_evaluations_ext_vec(dev, kl, synth.fill_ext(0xC0FFEE ^ 2, nl, start=rank * nl))
eq_lo = cb.build_eq_x_r_vec(dev, w[:2 * kl])
EQ = cb.wit_infer_by_monomial_expr(dev, [eq_lo], [(list(eq_slice_scalar(w[2 * kl:], rank)), [0])], kl)


def xchg(blob):
    outs = [None] * world
    dist.all_gather_object(outs, blob)
    return outs


comm = cb.Comm(dev, rank, world, xchg, barrier=dist.barrier)
terms = [([1, 0], [0, 1, 2])]
for dc in (True, False):
    prof, wall = [], []
    for i in range(8):
        dist.barrier()
        t0 = time.perf_counter()
        cb.prove_sharded(dev, comm, [EQ, A, B], terms, k, 3, cb.StandInTranscript(b"bench"), device_challenger=dc, flags=4)
        wall.append(time.perf_counter() - t0)
        if i >= 3:
            prof.append(dev.profile_last())
    p = np.mean(np.array(prof), axis=0)
    if rank == 0:
        print("device_challenger" if dc else "host", "wall_ms %.3f" % (1e3 * np.mean(wall[3:])), "sum_rounds %.3f" % p.sum(), [round(float(x), 4) for x in p])
comm.close()
dev.close()
dist.destroy_process_group()
