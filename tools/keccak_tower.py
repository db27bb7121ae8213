"""BASELINE config #4 shape on one GPU: the lookup tower of the keccak-f chip — 1094 lookup records per row
(ceno_zkvm/src/precompiles/lookup_keccakf.rs:97-101) over 2^rows_log rows — built and proven over VIRTUAL leaves
(cg_tower_build_virtual): the interleaved leaf layer (2^11 x rows, x 2 limbs) is never allocated.
usage: python tools/keccak_tower.py [rows_log=17] [n_records=1094] [base|ext]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceno_b200 as cb
from ceno_b200 import synth

rows_log = int(sys.argv[1]) if len(sys.argv) > 1 else 17
n_rec = int(sys.argv[2]) if len(sys.argv) > 2 else 1094
is_ext = not (len(sys.argv) > 3 and sys.argv[3] == "base")
dev = cb.Device(0)
n = 1 << rows_log
es = 16 if is_ext else 8
# one big allocation, record i at offset i * n (what the reference keeps: column-major sub-ranges of one buffer)
big = dev.alloc(es * n * n_rec)
chunk = synth.fill_ext(77, n) if is_ext else synth.fill_base(77, n)
pin, pinp = dev.pinned(chunk.nbytes)
pin[:] = chunk
for i in range(n_rec):                       # same values in every record (content does not change the cost)
    dev.h2d(big.ptr + es * n * i, pinp, chunk.nbytes)
dev.sync()
recs = [cb.MultilinearExtension(dev, cb.DeviceBuffer(dev, big.ptr + es * n * i, es * n, owner=False), rows_log, is_ext) for i in range(n_rec)]
info0 = dev.info()


def run():
    tw = cb.TowerProver.from_records(dev, [cb.VirtualTowerSpec(recs, n, [12345, 678], True)])
    dev.sync()
    t1 = time.perf_counter()
    proof, point = tw.create_proof(cb.StandInTranscript(b"keccak"))
    dev.sync()
    t2 = time.perf_counter()
    tw.close()
    return t1, t2, proof, point


run()
t0 = time.perf_counter()
t1, t2, proof, point = run()
l2m = (n_rec - 1).bit_length()
leaf_ext = 4 * (1 << (l2m + rows_log - 1))
out = {"config": "BASELINE #4 shape: keccak-f lookup tower over virtual leaves", "rows_log": rows_log, "records": n_rec, "record_field": "ext" if is_ext else "base",
       "record_bytes": es * n * n_rec, "virtual_leaf_layer_ext_elements": leaf_ext, "virtual_leaf_layer_bytes_not_allocated": 16 * leaf_ext,
       "tower_layers": l2m + rows_log, "build_ms": (t1 - t0) * 1e3, "prove_ms": (t2 - t1) * 1e3,
       "leaf_elements_per_s": leaf_ext / ((t2 - t0)), "proof_u64_words": int(proof.size), "point_len": int(point.size // 2),
       "device_free_bytes_before_towers": int(info0["free"])}
print(json.dumps(out))
dev.close()
