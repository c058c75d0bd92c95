#!/bin/bash
# ncu --set full capture of the hot kernels (one evaluation's worth), config 3, B=4096 (NCU_B overrides).  Usage: tools/ncu_full.sh <tag> [kernel-regex] [count]
tag=${1:-x}; rx=${2:-'k_refine|k_rot_bp2|k_rot_deriv|k_rot_energy|k_rot_build|k_hbond_coverage|k_env_coverage|k_backbone_fused|k_protein_hbond|k_cache_check'}; cnt=${3:-24}
mkdir -p gpurun_out
ncu --set full --metrics sm__sass_thread_inst_executed_op_fadd_pred_on.sum,sm__sass_thread_inst_executed_op_fmul_pred_on.sum,sm__sass_thread_inst_executed_op_ffma_pred_on.sum --clock-control none --import-source on --kernel-name "regex:$rx" --launch-skip ${4:-48} --launch-count $cnt \
    -o gpurun_out/full_$tag -f python tools/profile_run.py ${NCU_B:-4096} 4 > gpurun_out/full_$tag.log 2>&1
tail -2 gpurun_out/full_$tag.log
ls -la gpurun_out/full_$tag.ncu-rep
