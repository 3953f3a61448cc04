#!/bin/bash
# GPU session r02_h (2 GPUs): one process per GPU -- peer-memory exchange vs ncclSend/ncclRecv, every overlap mode; bench at N = 2.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_h_smi.txt; nvidia-smi topo -m >> gpurun_out/r02_h_smi.txt 2>&1
EQD_VERBOSE=1 timeout 1500 python -m pytest tests/test_gpu_nccl.py -m gpu -q --timeout 600 -x > gpurun_out/r02_h_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r02_h_status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_h_bench_n2.json 2> gpurun_out/r02_h_bench_n2.log
echo "bench n2 rc=$?" >> gpurun_out/r02_h_status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --halo 0 --no-cpu-baseline > gpurun_out/r02_h_bench_n2_nccl.json 2> gpurun_out/r02_h_bench_n2_nccl.log
echo "bench n2 nccl rc=$?" >> gpurun_out/r02_h_status.txt
cat gpurun_out/r02_h_status.txt; tail -5 gpurun_out/r02_h_pytest.log; grep "ms/step\|e2e leg\|peer memory\|parity" gpurun_out/r02_h_bench_n2.log | cut -c1-420; grep "ms/step" gpurun_out/r02_h_bench_n2_nccl.log | cut -c1-420
