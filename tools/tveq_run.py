"""One tower proof whose large layers run the general split-eq round kernel (tveq_round_kernel): driver for ncu captures.
usage: python tools/tveq_run.py [logup_nv=22] [iterations=1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceno_b200 as cb
from ceno_b200 import synth

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 22
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = cb.Device(0)
leaves = [cb.MultilinearExtension.from_evaluations_ext_vec(dev, nv, synth.fill_ext(60 + z, 1 << nv)) for z in range(4)]
spec = cb.TowerProverSpec(leaves, nv, True)
for _ in range(iters):
    tw = cb.TowerProver(dev, [spec])
    proof, point = tw.create_proof(cb.StandInTranscript(b"tveq"))
    tw.close()
print("ok", int(proof[0]))
dev.close()
