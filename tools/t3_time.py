"""Per-round device timing of T3-k (device challenger, CG_SC_PROFILE).
usage: python tools/t3_time.py [k] [reps] [table|virtual]   (virtual: eq handed over as its point, split-eq rounds)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceno_b200 as cb
from ceno_b200 import synth

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = cb.Device(0)
n = 1 << k
a = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, synth.fill_ext(0xC0FFEE ^ 1, n))
b = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, synth.fill_ext(0xC0FFEE ^ 2, n))
virt = len(sys.argv) > 3 and sys.argv[3] == "virtual"
eq = cb.EqPolynomial(dev, synth.fill_ext(0xE9, k)) if virt else cb.build_eq_x_r_vec(dev, synth.fill_ext(0xE9, k))
prof = []
for i in range(reps + 3):
    cb.IOPProverState.prove(dev, [eq, a, b], [([1, 0], [0, 1, 2])], k, 3, transcript=cb.StandInTranscript(b"bench"), device_challenger=True, flags=4)
    if i >= 3:
        prof.append(dev.profile_last())
p = np.mean(np.array(prof), axis=0)
print("virtual" if virt else "table", "tma", os.environ.get("CG_VEQ_TMA", "1"), "cfg", os.environ.get("CG_TOWER_CFG", "0"), "total_ms %.4f" % p.sum(), "r0 %.4f r1 %.4f r2 %.4f r3 %.4f" % tuple(p[:4]), "tail(sum r10..) %.4f" % p[10:].sum(), "rounds", np.round(p[:8], 4).tolist())
dev.close()
