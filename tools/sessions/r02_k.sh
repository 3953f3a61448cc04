#!/bin/bash
# GPU session r02_k (2 GPUs): eqd_set_host_comm on hardware -- the multi-process tests, bench at N = 2 (e2e without an
# NCCL communicator of the library's own), overlap 1 against the default 2.
mkdir -p gpurun_out
S=gpurun_out/r02_k_status.txt; : > $S
EQD_VERBOSE=1 timeout 1200 python -m pytest tests/test_gpu_nccl.py -m gpu -q --timeout 600 -x > gpurun_out/r02_k_pytest.log 2>&1; echo "pytest rc=$?" >> $S
run() { local name=$1 port=$2; shift 2
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 "$@" \
    > gpurun_out/r02_k_${name}.json 2> gpurun_out/r02_k_${name}.log; echo "$name rc=$?" >> $S; }
EQD_VERBOSE=1 run n2 29521 --steps 30 --warmup 5 --parity-steps 25
run n2_ov1 29522 --steps 30 --warmup 5 --overlap 1 --no-cpu-baseline
# one GPU each, side by side: L2 -> DRAM fetch size 32 B against the default (the list-following node kernels read single values a lattice row apart)
(CUDA_VISIBLE_DEVICES=0 EQD_VERBOSE=1 EQD_L2_FETCH=0 timeout 400 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r02_k_l2default.json 2> gpurun_out/r02_k_l2default.log; echo "l2default rc=$?" >> $S) &
(CUDA_VISIBLE_DEVICES=1 EQD_VERBOSE=1 EQD_L2_FETCH=32 timeout 400 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r02_k_l2f32.json 2> gpurun_out/r02_k_l2f32.log; echo "l2f32 rc=$?" >> $S) &
wait
cat $S; tail -4 gpurun_out/r02_k_pytest.log
for n in l2default l2f32; do echo "== $n"; grep "ms/step\|L2 fetch" gpurun_out/r02_k_${n}.log | tail -2 | cut -c1-400; done
for n in n2 n2_ov1; do echo "== $n"; grep "ms/step\|e2e leg\|parity over\|uploads" gpurun_out/r02_k_${n}.log | cut -c1-400; done
