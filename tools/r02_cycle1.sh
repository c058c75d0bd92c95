#!/bin/bash
# round-2 cycle 1: parity of the fast rotamer build + A/B timing against the Verlet/refine/prep path
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_node or golden or trajectory or full_batch or param_deriv" ) > gpurun_out/pytest_c1.log 2>&1
tail -8 gpurun_out/pytest_c1.log
python tools/quick_probe.py fast_build 2>&1 | tail -3
UPSIDE_B200_NO_FAST_BUILD=1 python tools/quick_probe.py verlet_prep 2>&1 | tail -3
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitize_c1.log python tools/sanitize_case.py > gpurun_out/sanitize_c1.out 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitize_c1.log
