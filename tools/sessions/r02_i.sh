#!/bin/bash
# GPU session r02_i (8 GPUs): the scaling line of the headline workload and the other configurations of BASELINE.json.
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/r02_i_host.txt; nproc >> gpurun_out/r02_i_host.txt
run() {  # name, port, args...
  local name=$1 port=$2; shift 2
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@" \
    > gpurun_out/r02_i_${name}.json 2> gpurun_out/r02_i_${name}.log
  echo "$name rc=$?" >> gpurun_out/r02_i_status.txt
}
: > gpurun_out/r02_i_status.txt
EQD_VERBOSE=1 run tpv104_n8 29601 --steps 20 --warmup 5
run tpv104_n8_long 29602 --steps 200 --warmup 5 --no-cpu-baseline
run tpv36_100m_n8 29603 --steps 50 --warmup 5 --case bench.tpv36_100m --decomp 2x2x2 --no-cpu-baseline
run tpv10_n8 29605 --steps 40 --warmup 5 --case test.tpv10 --decomp 2x2x2
run drva6_n8 29606 --steps 20 --warmup 5 --case test.drv.a6 --decomp 2x2x2
cat gpurun_out/r02_i_status.txt
for n in tpv104_n8 tpv104_n8_long tpv36_100m_n8 tpv10_n8 drva6_n8; do echo "== $n"; grep "ms/step\|e2e leg\|parity over" gpurun_out/r02_i_${n}.log | head -3 | cut -c1-330; done
