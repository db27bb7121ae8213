"""One RS-encode + Merkle commit of 64 columns x 2^20 rows (BASELINE config #5 shape) — the target of the ncu captures
of ntt_pass_kernel / p2_leaf_kernel.   usage: python tools/commit_run.py [log_n] [width] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceno_b200 as cb
from ceno_b200 import api, synth

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
width = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = cb.Device(0)
vals = synth.fill_base(0x9052, 8 * 8 + 22 + 8)
api.poseidon2_set_params(dev, vals[:64].reshape(8, 8), vals[64:86], vals[86:94], 0)
mat = dev.to_device(synth.fill_base(4242, width << log_n))
for _ in range(reps):
    code, tree, root = api.basefold_style_commit(dev, mat, width, log_n, 1)
    code.free()
    tree.free()
print("root", [hex(int(x)) for x in root])
dev.close()
