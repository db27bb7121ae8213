"""Per-round device timing of the sharded T3-k run (launch with torch.distributed.run).
usage: ... tools/t3_time_dist.py [k] [table|virtual] [extra_flags]"""
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
mode = sys.argv[2] if len(sys.argv) > 2 else "virtual"
xflags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if mode == "table":
    eq_lo = cb.build_eq_x_r_vec(dev, w[:2 * kl])
    EQ = cb.wit_infer_by_monomial_expr(dev, [eq_lo], [(list(eq_slice_scalar(w[2 * kl:], rank)), [0])], kl)
else:
    EQ = cb.EqPolynomial(dev, w, num_vars=kl)


def xchg(blob):
    outs = [None] * world
    dist.all_gather_object(outs, blob)
    return outs


comm = cb.Comm(dev, rank, world, xchg, barrier=dist.barrier)
terms = [([1, 0], [0, 1, 2])]
use_stream = os.environ.get("T3_STREAM", "0") == "1"
no_barrier = os.environ.get("T3_NOBARRIER", "0") == "1"
pf = 0 if os.environ.get("T3_NOPROF", "0") == "1" else 4
ts = torch.cuda.Stream() if use_stream else None
sh = ts.cuda_stream if use_stream else None
for dc in (True, False):
    prof, wall = [], []
    for i in range(8):
        if not no_barrier:
            dist.barrier()
        t0 = time.perf_counter()
        cb.prove_sharded(dev, comm, [EQ, A, B], terms, k, 3, cb.StandInTranscript(b"bench"), device_challenger=dc, flags=pf | xflags, stream=sh)
        wall.append(time.perf_counter() - t0)
        if i >= 3:
            prof.append(dev.profile_last() if pf else np.zeros(1))
    p = np.mean(np.array(prof), axis=0)
    if rank == 0:
        print(mode, "flags", xflags, "stream", use_stream, "nobarrier", no_barrier, "noprof", not pf, "device_challenger" if dc else "host", "wall_ms %.3f" % (1e3 * np.mean(wall[3:])), "sum_rounds %.3f" % p.sum(), [round(float(x), 4) for x in p])
comm.close()
dev.close()
dist.destroy_process_group()
