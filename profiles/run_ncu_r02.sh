#!/bin/bash
# profiles/run_ncu_r02.sh <tag> -- run on the GPU box (gpurun, ONE GPU): 
#   1. launch list (device time of every launch) of the default bench command
#   2. one `--set full --import-source on` capture (-lineinfo) of the hot kernels of one ADMM iteration on
#      (a) the whole 1024-problem batch, (b) the shard one of eight GPUs gets of it, (c) the forest scene
# The raw metric pages and the per-source-line stall tables (profiles/hotlines.py) are exported on the box; the .ncu-rep
# files stay in /tmp there (gpurun brings back at most 64 MiB).  Summarise here with
#   python profiles/summarize.py launches gpurun_out/launches_<tag>.csv profiles/r02_launches_batch.txt
#   python profiles/summarize.py rawcsv gpurun_out/raw_batch1024_<tag>.csv profiles/r02_full_batch1024.txt batch1024   (etc.)
tag=${1:-r02}
mkdir -p gpurun_out
export TRAJOPT_B200_NO_GRAPH=1
K='k_narrow|k_row_energy|k_row_grad|k_bp_count|k_bp_fill|k_bp_ccd|k_pack|k_piece|k_solve_bcr|k_slack|k_robot_ls|k_rows|k_en_items'
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_launch_${tag}.log 2>&1
# (a) whole batch: skip the build + warm-up iterations (about 30 kernels each), capture one iteration's worth
ncu --set full --clock-control none --import-source on --launch-skip 160 -c 32 -k regex:"$K" \
    -o /tmp/full_batch1024_${tag} -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_batch1024_${tag}.log 2>&1
ncu -i /tmp/full_batch1024_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_batch1024_${tag}.csv 2>/dev/null
for k in k_narrow k_row_energy k_row_grad; do python profiles/hotlines.py /tmp/full_batch1024_${tag}.ncu-rep $k 40 > gpurun_out/hot_batch1024_${k}_${tag}.txt 2>&1; done
# (b) batch shard 0/8
ncu --set full --clock-control none --import-source on --launch-skip 160 -c 32 -k regex:"$K" \
    -o /tmp/full_batch128_${tag} -f python bench.py --emulate-rank 0/8 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_batch128_${tag}.log 2>&1
ncu -i /tmp/full_batch128_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_batch128_${tag}.csv 2>/dev/null
# (c) forest
ncu --set full --clock-control none --import-source on --launch-skip 130 -c 30 -k regex:"$K" \
    -o /tmp/full_forest_${tag} -f python bench.py --workload forest --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_forest_${tag}.log 2>&1
ncu -i /tmp/full_forest_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_forest_${tag}.csv 2>/dev/null
for k in k_row_energy k_solve_bcr k_piece; do python profiles/hotlines.py /tmp/full_forest_${tag}.ncu-rep $k 30 > gpurun_out/hot_forest_${k}_${tag}.txt 2>&1; done
du -sh gpurun_out
