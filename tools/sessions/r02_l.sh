#!/bin/bash
# GPU session r02_l (8 GPUs): the 8-GPU line with the host communicator (e2e), overlap 1 against 2, drv.a6 against the
# oracle on the same decomposition.
mkdir -p gpurun_out
S=gpurun_out/r02_l_status.txt; : > $S
run() { local name=$1 port=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@" \
    > gpurun_out/r02_l_${name}.json 2> gpurun_out/r02_l_${name}.log; echo "$name rc=$?" >> $S; }
EQD_VERBOSE=1 run n8 29621 --steps 50 --warmup 5 --parity-steps 20
run n8_ov1 29622 --steps 50 --warmup 5 --overlap 1 --no-cpu-baseline
run drva6_n8 29623 --steps 20 --warmup 5 --case test.drv.a6 --decomp 2x2x2
cat $S
for n in n8 n8_ov1 drva6_n8; do echo "== $n"; grep "ms/step\|e2e leg\|parity over\|uploads" gpurun_out/r02_l_${n}.log | head -12 | cut -c1-400; done
