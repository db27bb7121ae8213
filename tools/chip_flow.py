"""Synthetic shard: several chips of BASELINE config #3 shape (2^20 cycles spread over ~6 opcode chips, num_vars 16..19)
proved end to end on the device — records, towers, tower proofs, main zerochecks, commitments — sequentially and on
ChipScheduler lanes.   usage: python tools/chip_flow.py [lanes...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ceno_b200 as cb
from ceno_b200 import api, synth
from ceno_b200 import chip as chipmod

dev = cb.Device(0)
vals = synth.fill_base(0x9052, 8 * 8 + 22 + 8)
api.poseidon2_set_params(dev, vals[:64].reshape(8, 8), vals[64:86], vals[86:94], 0)
shape = [(19, None), (19, 400000), (18, None), (18, 200000), (17, None), (16, 50000)]
if os.environ.get("CHIP_SHAPE"):      # e.g. CHIP_SHAPE=19:0,18:200000  (0 = full)
    shape = [(int(a.split(":")[0]), int(a.split(":")[1]) or None) for a in os.environ["CHIP_SHAPE"].split(",")]
COMMIT = os.environ.get("CHIP_COMMIT", "1") == "1"
chips = [chipmod.SyntheticChip(100 + s, k, ni, n_wit=24, n_read=4, n_write=4, n_lk=8) for s, (k, ni) in enumerate(shape)]
wits = [chipmod.upload_witness(dev, c) for c in chips]
tasks = [cb.ChipTask(i, c.estimated_memory_bytes(), payload=i, circuit_name=c.name) for i, c in enumerate(chips)]
rows = sum(c.num_instances for c in chips)


def prove(task, lane, stream):
    i = task.payload
    return chipmod.create_chip_proof(dev, chips[i], wits[i][1], cb.StandInTranscript(chips[i].name.encode()), stream=stream, commit_matrix=wits[i][0] if COMMIT else None)


out = {"chips": [{"num_vars": c.num_vars, "num_instances": c.num_instances, "witness_columns": c.n_wit,
                  "records": len(c.read_exprs) + len(c.write_exprs) + len(c.lk_exprs)} for c in chips], "rows": rows}
ref = None
for lanes in [int(a) for a in sys.argv[1:]] or [1, 4, 8]:
    cb.ChipScheduler(dev).execute(tasks, prove, lanes=lanes)       # warm-up (pools, lazy tables)
    l0 = dev.launch_count()
    t0 = time.perf_counter()
    res, tel = cb.ChipScheduler(dev).execute(tasks, prove, lanes=lanes)
    ms = (time.perf_counter() - t0) * 1e3
    if ref is None:
        ref = res
    same = all((a[k] == b[k]).all() for a, b in zip(ref, res) for k in a if not k.startswith("_"))
    out[f"lanes_{lanes}"] = {"ms": ms, "rows_per_s": rows / (ms * 1e-3), "launches": dev.launch_count() - l0, "identical_to_first": bool(same),
                             "per_chip_ms": [round(t["host_execution_ms"] + t["event_wait_ms"], 2) for t in tel]}
    if "_trace" in res[0]:
        out[f"lanes_{lanes}"]["trace"] = [r["_trace"] for r in res]
print(json.dumps(out))
dev.close()
