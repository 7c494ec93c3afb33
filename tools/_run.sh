set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed or 1080p or ncc" 2>&1 | tail -5) > gpurun_out/plan_tests.log 2>&1
for b in 256 128 64 32; do for sp in 0 1; do VBS_SEG_PLAN=$sp timeout 300 python tools/tc_probe.py $b 5 2>&1 | grep -E "^idp" | sed "s/^/overlap plan $sp: /"; done; done > gpurun_out/plan_times.log 2>&1
cat gpurun_out/plan_tests.log gpurun_out/plan_times.log
