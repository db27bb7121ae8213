"""Per-layer host-side time of the TOWER bench row (CG_TOWER_TRACE=1 prints eq / create / run / destroy per layer to stderr).
usage: CG_TOWER_TRACE=1 python tools/tower_trace.py [nvp=22]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceno_b200 as cb
from ceno_b200 import synth

nvp = int(sys.argv[1]) if len(sys.argv) > 1 else 22
nvl = nvp - 1
dev = cb.Device(0)
specs = []
for s in range(2):
    specs.append(cb.TowerProverSpec([cb.MultilinearExtension.from_evaluations_ext_vec(dev, nvp - 1, synth.fill_ext(50 + 2 * s + z, 1 << (nvp - 1))) for z in range(2)], nvp, False))
specs.append(cb.TowerProverSpec([None, None] + [cb.MultilinearExtension.from_evaluations_ext_vec(dev, nvl, synth.fill_ext(60 + z, 1 << nvl)) for z in range(2)], nvl, True))
for it in range(3):
    tw = cb.TowerProver(dev, specs)
    dev.sync()
    t0 = time.perf_counter()
    tw.create_proof(cb.StandInTranscript(b"tower"))
    dev.sync()
    print("prove_ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
    tw.close()
dev.close()
