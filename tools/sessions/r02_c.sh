#!/bin/bash
# GPU session r02_c: marching kernel v2 (thread = node, 8 x 16 node planes): parity tests, bench, --set full of k_march.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_march.py tests/test_gpu_branches.py -m gpu -q --timeout 400 > gpurun_out/r02_c_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r02_c_status.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c_bench.json 2> gpurun_out/r02_c_bench.log
echo "bench rc=$?" >> gpurun_out/r02_c_status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_march$|k_node_update3' \
  -s 4 -c 3 -o gpurun_out/r02_c_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c_full_bench.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r02_c_status.txt
ncu -i gpurun_out/r02_c_full.ncu-rep --page raw --csv > gpurun_out/r02_c_full_raw.csv 2>/dev/null
cat gpurun_out/r02_c_status.txt; tail -5 gpurun_out/r02_c_pytest.log; tail -3 gpurun_out/r02_c_bench.log | cut -c1-400
