"""Generate the `oracle_*.json` regression fixtures under tests/golden/ from the CPU restatement (schema: tests/golden/SCHEMA.md).
usage: python tools/gen_oracle_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402


def t3_instance(seed, k):
    n = 1 << k
    return (orc.fill_ext(0xE9 ^ (seed << 8), k), orc.fill_ext((0xC0FFEE ^ 1) ^ (seed << 32), n), orc.fill_ext((0xC0FFEE ^ 2) ^ (seed << 32), n))


def pairs(a):
    return [[int(x), int(y)] for x, y in a.reshape(-1, 2)]


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for seed in (0, 1):
        for k in (4, 10, 14):
            w, a, b = t3_instance(seed, k)
            eq = orc.build_eq_x_r_vec(w)
            rounds, fin, chal = orc.sumcheck_prove([(eq, True, k), (a, True, k), (b, True, k)], [([1, 0], [0, 1, 2])], k, 3, transcript=orc.Transcript(b"parity"))
            v = {"schema": "ceno_b200/t3/1", "generator": "oracle (stand-in sponge) — NOT an upstream vector", "seed": seed, "k": k, "degree": 3,
                 "transcript_label": "parity", "point_w": pairs(w), "eq_table_first8": pairs(eq[:16]),
                 "round_evaluations": [pairs(r) for r in rounds], "challenges": pairs(chal), "final_evaluations": pairs(fin)}
            with open(os.path.join(out, f"oracle_t3_seed{seed}_k{k}.json"), "w") as f:
                json.dump(v, f)
            print("wrote", f.name)


if __name__ == "__main__":
    main()
