"""The tower case of __graft_entry__.smoke() alone (split-eq layers forced at 2^10 points), checked against the oracle.
Used to confirm that the case also passes under a profiler that blocks the host in every launch (one launch per round):
  ncu --metrics gpu__time_duration.sum --clock-control none python tools/smoke_tower_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CG_TOWER_VEQ_MIN_NV"] = "10"
import ceno_b200 as cb
from oracle import oracle as orc

dev = cb.Device(0)
f1, f2 = orc.fill_ext(501, 1 << 13), orc.fill_ext(502, 1 << 13)
lq = [orc.fill_ext(503 + z, 1 << 13) for z in range(4)]
o_prod = [(orc.infer_tower_product_witness(14, f1, f2)[0], 14)]
o_lk = [(orc.infer_tower_logup_witness(13, *lq)[0], 14)]
want_proof, want_point = orc.tower_create_proof(o_prod, o_lk, orc.Transcript(b"smoke-tower"))
mk = lambda nv, x: cb.MultilinearExtension.from_evaluations_ext_vec(dev, nv, x)   # noqa: E731
tw = cb.TowerProver(dev, [cb.TowerProverSpec([mk(13, f1), mk(13, f2)], 14, False), cb.TowerProverSpec([mk(13, x) for x in lq], 13, True)])
got_proof, got_point = tw.create_proof(cb.StandInTranscript(b"smoke-tower"))
ok = np.array_equal(got_proof, want_proof) and np.array_equal(got_point, want_point)
tw.close()
print("tower case:", "bit-exact" if ok else "MISMATCH", "launches", dev.launch_count())
dev.close()
sys.exit(0 if ok else 1)
