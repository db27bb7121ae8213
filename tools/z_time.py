"""Per-round device time of the zerocheck-shaped generic instance Z (see tools/bench_rows.py)."""
import os, random, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceno_b200 as cb
from ceno_b200 import synth
P = 0xFFFFFFFF00000001
dev = cb.Device(0)
k, nb, ne, nt, deg = 20, 64, 4, 200, 4
n = 1 << k
rng = random.Random(1234)
mles = [cb.MultilinearExtension.from_evaluations_vec(dev, k, synth.fill_base(100 + i, n)) for i in range(nb)]
mles += [cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, synth.fill_ext(900 + i, n)) for i in range(ne)]
terms = []
for t in range(nt):
    sel = nb + rng.randrange(ne)
    wit = [rng.randrange(nb) for _ in range(rng.randint(1, deg - 1))]
    terms.append(([rng.randrange(P), rng.randrange(P)], [sel] + wit))
for flags in (4, 4 | 16):
    for _ in range(3):
        cb.IOPProverState.prove(dev, mles, terms, k, deg, transcript=cb.StandInTranscript(b"z"), flags=flags, device_challenger=True)
    p = dev.profile_last()
    print("plan" if flags == 4 else "noplan", "sum %.3f ms" % p.sum(), [round(float(x), 3) for x in p])
dev.close()
