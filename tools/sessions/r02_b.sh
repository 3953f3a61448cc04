#!/bin/bash
# GPU session r02_b: first run of the marching kernel -- triage, parity tests, bench line, launch list, --set full.
mkdir -p gpurun_out
timeout 300 python tools/march_debug.py test.tpv8 > gpurun_out/r02_b_debug.log 2>&1
echo "debug rc=$?" > gpurun_out/r02_b_status.txt
timeout 1200 python -m pytest tests/test_gpu_march.py -m gpu -q --timeout 400 > gpurun_out/r02_b_pytest_march.log 2>&1
echo "pytest_march rc=$?" >> gpurun_out/r02_b_status.txt
timeout 900 python -m pytest tests/test_gpu_branches.py -m gpu -q --timeout 400 > gpurun_out/r02_b_pytest_branches.log 2>&1
echo "pytest_branches rc=$?" >> gpurun_out/r02_b_status.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_b_bench.json 2> gpurun_out/r02_b_bench.log
echo "bench rc=$?" >> gpurun_out/r02_b_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_b_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_b_launches_bench.log 2>&1
echo "launches rc=$?" >> gpurun_out/r02_b_status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_march|k_tile_pml|k_node_update3|k_node_update12' \
  -s 8 -c 8 -o gpurun_out/r02_b_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_b_full_bench.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r02_b_status.txt
ncu -i gpurun_out/r02_b_full.ncu-rep --page raw --csv > gpurun_out/r02_b_full_raw.csv 2>/dev/null
cat gpurun_out/r02_b_status.txt; cat gpurun_out/r02_b_debug.log | tail -12; tail -5 gpurun_out/r02_b_pytest_march.log; tail -5 gpurun_out/r02_b_pytest_branches.log
tail -4 gpurun_out/r02_b_bench.log
