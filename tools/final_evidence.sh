#!/bin/bash
# Round-end evidence on ONE B200 (outputs under gpurun_out/, kept well below the 64 MiB copy-back limit: ncu reports are
# exported to raw CSV on the box and deleted).  usage: bash tools/final_evidence.sh [pytest]
O=gpurun_out
mkdir -p $O
rm -f $O/*.ncu-rep
if [ "${1:-}" = "pytest" ]; then timeout 600 python -m pytest tests -m gpu -x -q > $O/r2z_pytest_gpu.txt 2>&1; tail -3 $O/r2z_pytest_gpu.txt; fi
timeout 400 python bench.py --steps 10 --warmup 3 > $O/r2z_bench_n1.json 2> $O/r2z_bench_n1.err; tail -2 $O/r2z_bench_n1.err
timeout 200 python bench.py --steps 10 --warmup 3 --k 20 --cpu-k 20 --no-rows > $O/r2z_bench_k20.json 2> $O/r2z_bench_k20.err; tail -1 $O/r2z_bench_k20.err
timeout 100 python tools/keccak_tower.py 19 1094 > $O/r2z_keccak19.json 2>/dev/null
timeout 150 python tools/chip_flow.py 1 4 8 > $O/r2z_chip_flow.json 2> $O/r2z_chip_flow.err; tail -5 $O/r2z_chip_flow.err | cut -c1-300
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2z_launches_t3_24_virtual_devicemode.csv python tools/t3_run.py 24 2 device virtual > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2z_launches_smoke_under_ncu.csv python -c "import __graft_entry__ as g; g.smoke()" > $O/r2z_smoke_ncu.txt 2>&1; tail -1 $O/r2z_smoke_ncu.txt
timeout 300 ncu --set full --clock-control none -k regex:"veq_persist_kernel|veq_tma_kernel" -c 2 -f -o /tmp/r2z_veq python tools/t3_run.py 24 1 device virtual > /dev/null 2>&1
ncu -i /tmp/r2z_veq.ncu-rep --page raw --csv > $O/r2z_veq_raw.csv 2>/dev/null
# only the 2^24 leaf layer runs split-eq rounds: the first four launches are its rounds 0 .. 3
CG_TOWER_VEQ_MIN_NV=24 timeout 400 ncu --set full --clock-control none -k regex:"tveq_round_kernel" -c 4 -f -o /tmp/r2z_tveq python tools/tveq_run.py 24 1 > /dev/null 2>&1
ncu -i /tmp/r2z_tveq.ncu-rep --page raw --csv > $O/r2z_tveq_raw.csv 2>/dev/null
du -sh $O; ls -la $O | grep r2z
