"""Turn an Nsight Compute report into the trimmed per-launch summary committed under profiles/ (and, optionally, the
per-launch DRAM byte counts bench.py reports as `roofline.traffic`).

  ncu --set full --clock-control none --import-source on -k regex:<kernel> -c N -o gpurun_out/prof python tools/t3_run.py ...
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_<kernel>_ncu_summary.csv [--traffic KEY K]

--traffic KEY K : also write/refresh profiles/traffic.json[KEY] = {"k": K, "bytes_per_launch": [dram read + write per launch]}
Runs `ncu -i <report> --page raw --csv` (ncu is in the image; reading a report needs no GPU)."""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(value.replace(",", "")) * scale


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):   # already exported on the GPU box (`ncu -i rep --page raw --csv`; reports can exceed the copy-back limit)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    keep = [i for i, h in enumerate(hdr) if any(h == k or h.startswith(k) for k in KEEP) and ".min" not in h and ".max.pct" not in h]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch {i}" for i in range(len(launches))])
        for i in keep:
            w.writerow([hdr[i], units[i]] + [r[i] for r in launches])
    print(f"{out}: {len(keep)} metrics x {len(launches)} launches")
    if "--traffic" in sys.argv:
        key, k = sys.argv[sys.argv.index("--traffic") + 1], int(sys.argv[sys.argv.index("--traffic") + 2])
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        per = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in launches]
        path = os.path.join(os.path.dirname(os.path.abspath(out)), "traffic.json")
        tj = json.load(open(path)) if os.path.exists(path) else {}
        tj[key] = {"k": k, "bytes_per_launch": per, "source": os.path.basename(out)}
        json.dump(tj, open(path, "w"), indent=1)
        print(f"{path}[{key}] = {per}")


if __name__ == "__main__":
    main()
