#!/bin/bash
# GPU session r02_f (1 GPU): PML marching kernel -- parity, bench, launch list and --set full of the new kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_march.py -m gpu -q --timeout 500 > gpurun_out/r02_f_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r02_f_status.txt
EQD_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_f_bench.json 2> gpurun_out/r02_f_bench.log
echo "bench rc=$?" >> gpurun_out/r02_f_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_f_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02_f_launches.log 2>&1
echo "launches rc=$?" >> gpurun_out/r02_f_status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_march|k_node_update' \
  -s 3 -c 5 -o gpurun_out/r02_f_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_f_full_bench.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r02_f_status.txt
ncu -i gpurun_out/r02_f_full.ncu-rep --page raw --csv > gpurun_out/r02_f_full_raw.csv 2>/dev/null
cat gpurun_out/r02_f_status.txt; tail -5 gpurun_out/r02_f_pytest.log; grep "ms/step\|e2e leg" gpurun_out/r02_f_bench.log | cut -c1-500
