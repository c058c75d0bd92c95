#!/bin/bash
# round-2 development cycle: core parity tests + quick probe (+ optional extra command)
mkdir -p gpurun_out
tag=${1:-c}; shift
( time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_node or golden or trajectory or full_batch or param_deriv" ) > gpurun_out/pytest_$tag.log 2>&1
tail -4 gpurun_out/pytest_$tag.log
python tools/quick_probe.py $tag 2>&1 | tail -3
for c in "$@"; do bash -c "$c"; done
