#!/bin/bash
# Round-2 evidence for one build (run on the GPU box, results under gpurun_out/, copied into profiles/ by the caller):
#   launch list of tools/profile_run.py (ncu gpu__time_duration), ncu --set full (+ FP32 instruction counters) of the hot kernels,
#   traffic.json stamped with the source hash, bench line (N=1), compute-sanitizer memcheck of the small end-to-end case.
tag=${1:-r02x}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 220 --csv --log-file gpurun_out/${tag}_launches_config3_B4096.csv python tools/profile_run.py 4096 2 > gpurun_out/${tag}_prof.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches_config3_B4096.csv 40 > gpurun_out/${tag}_launch_summary.txt; head -12 gpurun_out/${tag}_launch_summary.txt
bash tools/ncu_full.sh $tag '' 40 44 > /dev/null 2>&1
python tools/ncu_table.py gpurun_out/full_$tag.ncu-rep > gpurun_out/${tag}_ncu_full_table.txt 2>&1
python tools/make_traffic.py gpurun_out/full_$tag.ncu-rep profiles/traffic.json $tag > gpurun_out/${tag}_traffic.log 2>&1; cp profiles/traffic.json gpurun_out/${tag}_traffic.json
[ -n "$KEEP_REP" ] || rm -f gpurun_out/full_$tag.ncu-rep   # (gpurun brings back at most 64 MiB)
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','us_per_force_eval')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])"
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/${tag}_sanitize_memcheck.log python tools/sanitize_case.py > /dev/null 2>&1; tail -1 gpurun_out/${tag}_sanitize_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --log-file gpurun_out/${tag}_sanitize_racecheck.log python tools/sanitize_case.py > /dev/null 2>&1; tail -1 gpurun_out/${tag}_sanitize_racecheck.log
