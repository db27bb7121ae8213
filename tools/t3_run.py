"""Minimal T3-k driver for ncu captures: a few sumcheck steps with resident inputs, nothing else.
usage: python tools/t3_run.py [k] [steps] [host|device] [table|virtual]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceno_b200 as cb
from ceno_b200 import synth

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "device"
dev = cb.Device(0)
n = 1 << k
a = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, synth.fill_ext(0xC0FFEE ^ 1, n))
b = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, synth.fill_ext(0xC0FFEE ^ 2, n))
virt = len(sys.argv) > 4 and sys.argv[4] == "virtual"
eq = cb.EqPolynomial(dev, synth.fill_ext(0xE9, k)) if virt else cb.build_eq_x_r_vec(dev, synth.fill_ext(0xE9, k))
for _ in range(steps):
    out = cb.IOPProverState.prove(dev, [eq, a, b], [([1, 0], [0, 1, 2])], k, 3, transcript=cb.StandInTranscript(b"bench"),
                                  device_challenger=(mode == "device"))
print("ok", out[0][0])
dev.close()
