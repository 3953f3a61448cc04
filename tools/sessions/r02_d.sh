#!/bin/bash
# GPU session r02_d: marching kernel with ghost sharing (march = 2, v/d double-buffered): parity, bench with set-up laps, ncu.
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/r02_d_host.txt; nproc >> gpurun_out/r02_d_host.txt; nvidia-smi -L >> gpurun_out/r02_d_host.txt
timeout 1500 python -m pytest tests/test_gpu_march.py tests/test_gpu_parity.py -m gpu -q --timeout 500 -x > gpurun_out/r02_d_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r02_d_status.txt
EQD_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_d_bench.json 2> gpurun_out/r02_d_bench.log
echo "bench rc=$?" >> gpurun_out/r02_d_status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_march$|k_node_update3' \
  -s 4 -c 3 -o gpurun_out/r02_d_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_d_full_bench.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r02_d_status.txt
ncu -i gpurun_out/r02_d_full.ncu-rep --page raw --csv > gpurun_out/r02_d_full_raw.csv 2>/dev/null
cat gpurun_out/r02_d_status.txt; tail -5 gpurun_out/r02_d_pytest.log; grep "ms/step\|e2e leg\|plan_march" gpurun_out/r02_d_bench.log | cut -c1-600
