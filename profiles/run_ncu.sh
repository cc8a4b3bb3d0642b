#!/bin/bash
# profiles/run_ncu.sh <tag> -- run on the GPU box (gpurun): launch list + one --set full capture of the hot kernels of one
# ADMM iteration on the bench workload.  Outputs land in gpurun_out/; summarise here with profiles/summarize.py.
tag=${1:-rXX}
mkdir -p gpurun_out
export TRAJOPT_B200_NO_GRAPH=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 150 -c 60 \
    -k regex:'k_narrow|k_solve_bcr|k_piece|k_row_energy|k_bp_count|k_bp_fill|k_bp_ccd|k_slack|k_robot_ls|k_row_grad|k_pack' \
    -o gpurun_out/full_${tag} -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full_${tag}.log 2>&1
ls -la gpurun_out/
