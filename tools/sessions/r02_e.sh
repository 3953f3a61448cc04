#!/bin/bash
# GPU session r02_e (1 GPU): march = 1 vs march = 2 in situ (packed element planes), set-up laps, launch list.
mkdir -p gpurun_out
EQD_VERBOSE=1 timeout 400 python bench.py --steps 40 --warmup 5 --march 1 --no-cpu-baseline > gpurun_out/r02_e_bench_m1.json 2> gpurun_out/r02_e_bench_m1.log
echo "bench m1 rc=$?" > gpurun_out/r02_e_status.txt
timeout 400 python bench.py --steps 40 --warmup 5 --march 2 --no-cpu-baseline > gpurun_out/r02_e_bench_m2.json 2> gpurun_out/r02_e_bench_m2.log
echo "bench m2 rc=$?" >> gpurun_out/r02_e_status.txt
timeout 400 python bench.py --steps 40 --warmup 5 --march 0 --no-cpu-baseline > gpurun_out/r02_e_bench_m0.json 2> gpurun_out/r02_e_bench_m0.log
echo "bench m0 rc=$?" >> gpurun_out/r02_e_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_e_launches_m2.csv \
  python bench.py --steps 4 --warmup 3 --march 2 --no-cpu-baseline > gpurun_out/r02_e_launches_m2.log 2>&1
echo "launches rc=$?" >> gpurun_out/r02_e_status.txt
timeout 900 python -m pytest tests/test_gpu_march.py -m gpu -q --timeout 500 -x > gpurun_out/r02_e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_e_status.txt
cat gpurun_out/r02_e_status.txt; grep "ms/step\|e2e leg" gpurun_out/r02_e_bench_m*.log | cut -c1-420; tail -3 gpurun_out/r02_e_pytest.log
