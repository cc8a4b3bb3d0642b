#!/bin/bash
# profiles/run_ncu_r02b.sh <tag> -- final captures of round 2; run on the GPU box (gpurun, ONE GPU):
#   1. launch list (device time of every launch of the ITERATION kernels; the 14 k build launches of the 1024 clouds are
#      filtered out by name) of the default bench command
#   2. one `--set full --import-source on` capture (-lineinfo) of the kernels of one ADMM iteration on
#      (a) the whole 1024-problem batch, (b) the shard one of eight GPUs gets of it, (c) the forest scene
# Raw metric pages and per-source-line stall tables are exported on the box (the .ncu-rep files stay in /tmp there).
# Summarise here with
#   python profiles/summarize.py launches gpurun_out/launches_<tag>.csv profiles/r02b_launches_batch.txt
#   python profiles/summarize.py rawcsv gpurun_out/raw_batch1024_<tag>.csv profiles/r02b_full_batch1024.txt batch1024   (etc.)
tag=${1:-r02b}
mkdir -p gpurun_out
export TRAJOPT_B200_NO_GRAPH=1
IT='^k_(rows|bp_count|bp_top|bp_fill|bp_ccd|narrow|np_top|pack|en_items|row_grad|row_energy|row_smax|robot_ls|piece|solve|ls_|apply_step|slack|self_|couple|flags)'
K='^k_(narrow|row_energy|row_grad|bp_count|bp_fill|bp_ccd|pack|piece|solve_bcr|slack|robot_ls|rows|ls_bound_mask|np_top|bp_top)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$IT" -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_launch_${tag}.log 2>&1
# (a) whole batch: skip the warm-up iterations (about 30 kernels each), capture one iteration's worth
timeout 1200 ncu --set full --clock-control none --import-source on --launch-skip 100 -c 34 -k regex:"$K" \
    -o /tmp/full_batch1024_${tag} -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_batch1024_${tag}.log 2>&1
ncu -i /tmp/full_batch1024_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_batch1024_${tag}.csv 2>/dev/null
for k in k_narrow k_row_energy k_row_grad k_bp_count; do python profiles/hotlines.py /tmp/full_batch1024_${tag}.ncu-rep $k 40 > gpurun_out/hot_batch1024_${k}_${tag}.txt 2>&1; done
# (b) batch shard 0/8
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 100 -c 34 -k regex:"$K" \
    -o /tmp/full_batch128_${tag} -f python bench.py --emulate-rank 0/8 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_batch128_${tag}.log 2>&1
ncu -i /tmp/full_batch128_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_batch128_${tag}.csv 2>/dev/null
# (c) forest
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 90 -c 30 -k regex:"$K" \
    -o /tmp/full_forest_${tag} -f python bench.py --workload forest --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_forest_${tag}.log 2>&1
ncu -i /tmp/full_forest_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_forest_${tag}.csv 2>/dev/null
for k in k_row_energy k_solve_bcr k_piece k_narrow; do python profiles/hotlines.py /tmp/full_forest_${tag}.ncu-rep $k 30 > gpurun_out/hot_forest_${k}_${tag}.txt 2>&1; done
du -sh gpurun_out
