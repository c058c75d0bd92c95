#!/bin/bash
# One GPU development cycle: parity tests, launch list (ncu, B=4096, 2 rounds), short bench.  Usage: tools/gpu_cycle.sh <tag> [pytest-args]
tag=${1:-x}; shift
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q "$@" ) > gpurun_out/pytest_$tag.log 2>&1
tail -5 gpurun_out/pytest_$tag.log
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 220 --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_run.py 4096 2 > gpurun_out/prof_$tag.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$tag.csv 30
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json'))
print({k:d[k] for k in ('value','ms_per_step','us_per_force_eval')}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'] and d['cpu_baseline']['value'])"
