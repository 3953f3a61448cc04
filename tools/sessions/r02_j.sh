#!/bin/bash
# GPU session r02_j (1 GPU): the whole -m gpu suite; the benchmark line (default invocation), ghost columns along the
# fast axis (march 3), the 1/8-size case (what one rank of an 8-GPU run holds), a 150-step line with parity past the
# first rupture; launch list and --set full of the step's kernels at benchmark size.
mkdir -p gpurun_out
S=gpurun_out/r02_j_status.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_j_pytest.log 2>&1; echo "pytest rc=$?" >> $S
EQD_VERBOSE=1 timeout 600 python bench.py > gpurun_out/r02_j_bench.json 2> gpurun_out/r02_j_bench.log; echo "bench rc=$?" >> $S
timeout 400 python bench.py --march 3 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_j_bench_m3.json 2> gpurun_out/r02_j_bench_m3.log; echo "bench m3 rc=$?" >> $S
timeout 400 python bench.py --case bench.tpv104_200m --steps 100 --warmup 5 --parity-steps 30 > gpurun_out/r02_j_bench_200m.json 2> gpurun_out/r02_j_bench_200m.log; echo "bench 200m rc=$?" >> $S
timeout 700 python bench.py --steps 150 --warmup 5 --parity-steps 155 > gpurun_out/r02_j_bench_150.json 2> gpurun_out/r02_j_bench_150.log; echo "bench 150 rc=$?" >> $S
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_j_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-process-warmup > gpurun_out/r02_j_launches.log 2>&1; echo "launches rc=$?" >> $S
timeout 800 ncu --set full --clock-control none --import-source on -k regex:'^k_march$|^k_march_pml$|k_node_update|k_assemble' \
  -s 35 -c 10 -o gpurun_out/r02_j_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-process-warmup > gpurun_out/r02_j_full_bench.log 2>&1; echo "ncu full rc=$?" >> $S
ncu -i gpurun_out/r02_j_full.ncu-rep --page raw --csv > gpurun_out/r02_j_full_raw.csv 2>/dev/null
[ $(stat -c%s gpurun_out/r02_j_full.ncu-rep) -gt 40000000 ] && rm -f gpurun_out/r02_j_full.ncu-rep
cat $S; tail -4 gpurun_out/r02_j_pytest.log
for n in bench bench_m3 bench_200m bench_150; do echo "== $n"; grep "ms/step\|e2e leg\|parity over" gpurun_out/r02_j_${n}.log | cut -c1-400; done
