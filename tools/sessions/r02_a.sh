#!/bin/bash
# GPU session r02_a (run through gpurun from the repo root): new branch parity tests, the bench line with
# its parity block, the launch list and the --set full capture of the four step kernels on the 100 m mesh,
# and the round-1 tile / bank_order variants that were never timed.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r02_a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_branches.py -m gpu -q -x > gpurun_out/r02_a_pytest_branches.log 2>&1
echo "pytest_branches rc=$?" >> gpurun_out/r02_a_status.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_a_bench.json 2> gpurun_out/r02_a_bench.log
echo "bench rc=$?" >> gpurun_out/r02_a_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_a_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_a_launches_bench.log 2>&1
echo "launches rc=$?" >> gpurun_out/r02_a_status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tile_reg|k_tile_pml|k_node_update3|k_node_update12|k_fault' \
  -s 10 -c 10 -o gpurun_out/r02_a_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_a_full_bench.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r02_a_status.txt
ncu -i gpurun_out/r02_a_full.ncu-rep --page raw --csv > gpurun_out/r02_a_full_raw.csv 2>/dev/null
EQD_TUNE_BANK_ORDER=2 timeout 600 python tools/tune_tiles.py > gpurun_out/r02_a_tune_bank2.log 2>&1
echo "tune bank2 rc=$?" >> gpurun_out/r02_a_status.txt
EQD_TUNE_BANK_ORDER=0 timeout 300 python tools/tune_tiles.py bench.tpv104_100m 2 > gpurun_out/r02_a_tune_bank0.log 2>&1
cat gpurun_out/r02_a_status.txt
tail -3 gpurun_out/r02_a_pytest_branches.log
cat gpurun_out/r02_a_bench.json | head -c 3000
