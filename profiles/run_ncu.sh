#!/bin/bash
# profiles/run_ncu.sh <tag> -- run on the GPU box (gpurun): launch list + one --set full capture of the hot kernels of one
# ADMM iteration on the bench workload (forest) and of the throughput-regime kernels on the 256-problem batch.
# gpurun copies back at most 64 MiB: the raw metric pages are exported to CSV on the box and only the (smaller) forest
# report is kept for the source-level view (profiles/hotlines.py).
tag=${1:-rXX}
mkdir -p gpurun_out
export TRAJOPT_B200_NO_GRAPH=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/ncu_launch_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 72 -c 24 \
    -k regex:'k_narrow|k_solve_bcr|k_piece|k_row_energy|k_bp_count|k_bp_fill|k_bp_ccd|k_slack|k_robot_ls|k_row_grad|k_pack|k_rows' \
    -o gpurun_out/full_${tag} -f python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/ncu_full_${tag}.log 2>&1
ncu -i gpurun_out/full_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_${tag}.csv 2>/dev/null
ncu --set full --clock-control none --launch-skip 18 -c 9 \
    -k regex:'k_narrow|k_row_energy|k_row_grad|k_bp_count|k_bp_fill|k_pack' \
    -o /tmp/full_batch_${tag} -f python bench.py --workload batch --problems 256 --steps 2 --warmup 3 > gpurun_out/ncu_full_batch_${tag}.log 2>&1
ncu -i /tmp/full_batch_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_batch_${tag}.csv 2>/dev/null   # summarise with: python profiles/summarize.py rawcsv ... batch256
du -sh gpurun_out
