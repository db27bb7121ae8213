"""Phase breakdown of the cluster tail kernel (CG_TAIL_DEBUG=1): one T3-k sumcheck that runs entirely in the tail."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CG_TAIL_DEBUG"] = "1"
import ceno_b200 as cb
from ceno_b200 import synth

dev = cb.Device(0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
devch = (sys.argv[2] == "dev") if len(sys.argv) > 2 else True
n = 1 << k
w = synth.fill_ext(5, k)
a = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, synth.fill_ext(6, n))
b = cb.MultilinearExtension.from_evaluations_ext_vec(dev, k, synth.fill_ext(7, n))
eq = cb.build_eq_x_r_vec(dev, w)
for it in range(2):
    t0 = time.perf_counter()
    cb.IOPProverState.prove(dev, [eq, a, b], [([1, 0], [0, 1, 2])], k, 3, transcript=cb.StandInTranscript(b"dbg"), device_challenger=devch)
    print("wall ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
