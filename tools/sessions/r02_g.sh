#!/bin/bash
# GPU session r02_g (1 GPU): PML bundles with turned node planes; bench; eqdyna_host -o; racecheck on wedges.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_march.py -m gpu -q --timeout 500 -x > gpurun_out/r02_g_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r02_g_status.txt
EQD_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_g_bench.json 2> gpurun_out/r02_g_bench.log
echo "bench rc=$?" >> gpurun_out/r02_g_status.txt
# stand-alone driver: writes the reference's output files; compare frt.txt* with the goldens (check.test.py criterion) in python
timeout 300 eqdyna_b200/bin/eqdyna_host tests/golden/cases/test.tpv8 -o gpurun_out/r02_g_host_tpv8 -march 1 -box 2 > gpurun_out/r02_g_host_tpv8.log 2>&1
echo "eqdyna_host rc=$?" >> gpurun_out/r02_g_status.txt
timeout 300 python - > gpurun_out/r02_g_host_check.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "tests")
import golden_io
for f in ("frt.txt0", "frt.txt2"):
    ok, msg = golden_io.compare_txt_files(golden_io.golden_path("test.tpv8", f), os.path.join("gpurun_out/r02_g_host_tpv8", f))
    print(f, ok, msg)
PY
echo "host check rc=$?" >> gpurun_out/r02_g_status.txt
# racecheck over the shared-memory assembly with wedge colours (tpv36, 3 steps, one sub-domain) and the marching kernels (tpv8)
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -c "
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import parity
w = parity.build_world('test.tpv37', (1,1,1), 3)
parity.run_gpu(w, options={'box': 2, 'box_compact': 1}, pre_options={'march': 1})
print('ran tpv37 3 steps under racecheck')
" > gpurun_out/r02_g_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_g_status.txt
cat gpurun_out/r02_g_status.txt; tail -3 gpurun_out/r02_g_pytest.log; grep "ms/step\|e2e leg" gpurun_out/r02_g_bench.log | cut -c1-500; cat gpurun_out/r02_g_host_check.log; tail -4 gpurun_out/r02_g_racecheck.log
