#!/bin/bash
# Runs on the GPU box (one gpurun call): ncu evidence for profiles/ (B=8 audio-visual, tools/ncu_one_eval.py =
# weight prep + conditioning + 2 denoiser evaluations).  Outputs land in gpurun_out/.
set -u
R=${ROUND:-r2}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
PY="python tools/ncu_one_eval.py"
# (1) launch list of the whole run
$NCU --metrics gpu__time_duration.sum -c 800 --csv --log-file gpurun_out/${R}_launches.csv $PY > /dev/null 2>&1
# (2) DRAM traffic + duration of every GEMM-class launch
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:'gemm_tc|mlp_fused|splitk_reduce' \
     --csv --log-file gpurun_out/${R}_gemm_traffic.csv $PY > /dev/null 2>&1
# (3) speed-of-light sections of the GEMM-class launches of the second evaluation
#     (4 conditioning GEMMs + 43 of the first evaluation are skipped; gemm_tc only: 39 per evaluation, mt_proj is #38, upembed.conv1 of stage 1 #19)
$NCU --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
     -k regex:'gemm_tc|mlp_fused' --launch-skip 47 --launch-count 43 -o gpurun_out/${R}_gemm_sol -f $PY > /dev/null 2>&1
# (4) full captures: mt_proj + fused head (last GEMM of the second evaluation) and upembed.conv1 of stage 1
$NCU --set full --import-source on -k regex:gemm_tc --launch-skip 81 --launch-count 1 -o gpurun_out/${R}_gemm_mtproj -f $PY > /dev/null 2>&1
$NCU --set full --import-source on -k regex:gemm_tc --launch-skip 62 --launch-count 1 -o gpurun_out/${R}_gemm_upconv1 -f $PY > /dev/null 2>&1
# (4b) full capture of the fused MLP chain of the last stage (last mlp_fused launch of the second evaluation)
$NCU --set full --import-source on -k regex:mlp_fused --launch-skip 7 --launch-count 1 -o gpurun_out/${R}_mlp_fused -f $PY > /dev/null 2>&1
# (5) memory-bound kernels of the second evaluation
$NCU --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SchedulerStats \
     -k regex:'q_dwln|qv_tile|pool_ln|av_gate|kpool|ms_sum|gn_apply|gn_stats|gn_fused|ln_vec|upsample2x|stem|temb|final_up|attn_fold|attn_operands|axpy' \
     -o gpurun_out/${R}_membound -f $PY > /dev/null 2>&1
ls -la gpurun_out | tail -12
