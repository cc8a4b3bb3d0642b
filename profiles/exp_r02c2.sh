#!/bin/bash
# GJK with the hull vertices read from L1 every round (register room -> more resident warps): tests, A/B; launch list of the line search
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_policy.py tests/test_gpu_parity.py tests/test_gpu_batch.py -x -q 2>&1 | tail -8 ) > gpurun_out/r02c2_tests.log
cat gpurun_out/r02c2_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02c2_$tag.json 2> gpurun_out/r02c2_$tag.err; }
run default A=1
run pmem4 TRAJOPT_B200_NP_PMEM=4
run pmem5 TRAJOPT_B200_NP_PMEM=5
run pmem6 TRAJOPT_B200_NP_PMEM=6
timeout 600 python bench.py --workload forest --steps 30 --warmup 5 --no-cpu > gpurun_out/r02c2_forest.json 2> gpurun_out/r02c2_forest.err
TRAJOPT_B200_NP_PMEM=5 timeout 600 python bench.py --workload forest --steps 30 --warmup 5 --no-cpu > gpurun_out/r02c2_forest_pmem5.json 2> gpurun_out/r02c2_forest_pmem5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02c2_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02c2_")[1][:-5].ljust(14), "ms/step %.3f"%j["ms_per_step"], "e2e %.0f"%j["e2e"]["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_row_energy","k_row_grad","k_bp_ccd","k_bp_count","k_bp_fill","k_pack","k_robot_ls","k_piece","k_solve_bcr","k_slack") if n in k))
    except Exception as e:
        print(f, "ERR", e)
PY
# launch list of two iterations of the shard (un-graphed): per-launch durations and grids of the line-search kernels
TRAJOPT_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -k regex:'^k_(row_energy|robot_ls|row_grad|ls_|narrow|bp_|pack|np_|piece|solve|slack|rows|en_items|apply)' --launch-skip 150 -c 120 --csv --log-file gpurun_out/r02c2_launches_shard.csv python bench.py --emulate-rank 0/8 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02c2_ncu.log 2>&1
tail -3 gpurun_out/r02c2_ncu.log
