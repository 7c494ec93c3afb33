set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core" 2>&1 | tail -15) > gpurun_out/tc_tests.log 2>&1
timeout 300 python tools/tc_probe.py 256 5 2>&1 | grep -E "^tc" > gpurun_out/tc_times.log 2>&1
VBS_TC_TIMELINE=1 timeout 300 python tools/tc_probe.py 64 1 2>&1 | grep -A40 "tc timeline" | head -30 >> gpurun_out/tc_times.log 2>&1
cat gpurun_out/tc_tests.log gpurun_out/tc_times.log
