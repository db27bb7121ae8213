"""BASELINE config #4 at full size: the keccak-f lookup tower (1094 records per row) over 2^rows_log rows SHARDED across the
GPUs of one box — every rank holds rows / N rows of every record (its rows of both fan-in blocks), virtual leaves, layer
kernels storing into the partner ranks' buffers over NVLink, big layers' sumchecks sharded.
launch: python -m torch.distributed.run --nproc-per-node N tools/keccak_tower_sharded.py [rows_log=22] [n_records=1094]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import ceno_b200 as cb
from ceno_b200 import synth

rows_log = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n_rec = int(sys.argv[2]) if len(sys.argv) > 2 else 1094
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = world.bit_length() - 1
dev = cb.Device(local)
n_loc = 1 << (rows_log - g)


def xchg(blob):
    outs = [None] * world
    dist.all_gather_object(outs, blob)
    return outs


comm = cb.Comm(dev, rank, world, xchg, barrier=dist.barrier)
l2m = (n_rec - 1).bit_length()
top_local = 1 << (l2m + rows_log - 1 - g)                 # entries of one leaf array on this rank
arena = 16 * 4 * (top_local + (1 << 22))                  # 4 arrays x (1/2 + 1/4 + ...) of the leaf slice, plus slack
comm.create_arena(arena, xchg)
big = dev.alloc(16 * n_loc * n_rec)
chunk = synth.fill_ext(77 + rank, n_loc)
pin, pinp = dev.pinned(chunk.nbytes)
pin[:] = chunk
for i in range(n_rec):
    dev.h2d(big.ptr + 16 * n_loc * i, pinp, chunk.nbytes)
dev.sync()
recs = [cb.MultilinearExtension(dev, cb.DeviceBuffer(dev, big.ptr + 16 * n_loc * i, 16 * n_loc, owner=False), rows_log - g, True) for i in range(n_rec)]


def run():
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    tw = cb.TowerProver.from_records(dev, [cb.VirtualTowerSpec(recs, n_loc, [12345, 678], True)], comm=comm)
    dev.sync()
    dist.barrier()
    t1 = time.perf_counter()
    proof, point = tw.create_proof(cb.StandInTranscript(b"keccak"))
    dev.sync()
    dist.barrier()
    t2 = time.perf_counter()
    tw.close()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, proof, point


run()
b_ms, p_ms, proof, point = run()
digest = int(np.bitwise_xor.reduce(proof))
alld = [None] * world
dist.all_gather_object(alld, digest)
if rank == 0:
    print(json.dumps({"config": "BASELINE #4: keccak-f lookup tower, rows sharded over the GPUs of one box, virtual leaves", "n_gpus": world,
                      "rows_log": rows_log, "rows_per_gpu_log": rows_log - g, "records": n_rec, "record_bytes_per_gpu": 16 * n_loc * n_rec,
                      "virtual_leaf_layer_ext_elements_global": 4 << (l2m + rows_log - 1), "tower_layers": l2m + rows_log,
                      "build_ms": b_ms, "prove_ms": p_ms, "proof_identical_on_all_ranks": len(set(alld)) == 1, "arena_bytes_per_gpu": arena,
                      "timing": "host wall clock between rank barriers (build includes the NVLink layer shuffles)"}))
comm.close()
dev.close()
dist.destroy_process_group()
