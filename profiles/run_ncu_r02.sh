#!/bin/bash
# profiles/run_ncu_r02.sh <tag> -- run on the GPU box (gpurun): one --set full capture (source-level, -lineinfo) of the hot
# kernels of an ADMM iteration on (a) the shard one of eight GPUs gets of the 1024-problem batch and (b) the forest scene,
# The raw metric pages and the per-source-line stall tables (profiles/hotlines.py) are exported on the box; the .ncu-rep
# files stay in /tmp there (gpurun brings back at most 64 MiB).
tag=${1:-r02}
mkdir -p gpurun_out
export TRAJOPT_B200_NO_GRAPH=1
K='k_narrow|k_row_energy|k_row_grad|k_bp_count|k_bp_fill|k_bp_ccd|k_pack|k_piece|k_solve_bcr|k_slack|k_robot_ls|k_rows'
# (a) batch shard 0/8: skip the warm-up iterations (about 22 kernels each), capture one iteration's worth
ncu --set full --clock-control none --import-source on --launch-skip 110 -c 26 -k regex:"$K" \
    -o /tmp/full_batch128_${tag} -f python bench.py --emulate-rank 0/8 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_batch128_${tag}.log 2>&1
ncu -i /tmp/full_batch128_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_batch128_${tag}.csv 2>/dev/null
# (b) forest
ncu --set full --clock-control none --import-source on --launch-skip 110 -c 26 -k regex:"$K" \
    -o /tmp/full_forest_${tag} -f python bench.py --workload forest --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_forest_${tag}.log 2>&1
ncu -i /tmp/full_forest_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_forest_${tag}.csv 2>/dev/null
for k in k_narrow k_row_energy k_row_grad k_bp_fill k_pack; do python profiles/hotlines.py /tmp/full_batch128_${tag}.ncu-rep $k 40 > gpurun_out/hot_batch128_${k}_${tag}.txt 2>&1; done
for k in k_row_energy k_solve_bcr k_piece k_narrow k_robot_ls; do python profiles/hotlines.py /tmp/full_forest_${tag}.ncu-rep $k 30 > gpurun_out/hot_forest_${k}_${tag}.txt 2>&1; done
du -sh gpurun_out
