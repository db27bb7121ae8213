"""Tower prover timing (2 product specs + 1 logup spec), optional per-layer host trace with CG_TOWER_TRACE=1.
usage: python tools/tower_time.py [nv_prod] [nv_logup] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceno_b200 as cb
from ceno_b200 import synth

nvp = int(sys.argv[1]) if len(sys.argv) > 1 else 22
nvl = int(sys.argv[2]) if len(sys.argv) > 2 else 21
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = cb.Device(0)
specs = []
for s in range(2):
    specs.append(cb.TowerProverSpec([cb.MultilinearExtension.from_evaluations_ext_vec(dev, nvp - 1, synth.fill_ext(50 + 2 * s + z, 1 << (nvp - 1))) for z in range(2)], nvp, False))
specs.append(cb.TowerProverSpec([None, None] + [cb.MultilinearExtension.from_evaluations_ext_vec(dev, nvl, synth.fill_ext(60 + z, 1 << nvl)) for z in range(2)], nvl, True))
ts = []
for i in range(reps + 2):
    dev.sync()
    t0 = time.perf_counter()
    tw = cb.TowerProver(dev, specs)
    t1 = time.perf_counter()
    tw.create_proof(cb.StandInTranscript(b"tower"))
    t2 = time.perf_counter()
    tw.close()
    if i >= 2:
        ts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
print("tower nv", nvp, nvl, "build_ms %.3f prove_ms %.3f" % (sum(t[0] for t in ts) / len(ts), sum(t[1] for t in ts) / len(ts)))
dev.close()
