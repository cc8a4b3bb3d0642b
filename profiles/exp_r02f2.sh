#!/bin/bash
# line search starts at the first rung that is not infeasible for certain (k_row_smax): whole suite, policies
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02f2_tests.log 2>&1
cat gpurun_out/r02f2_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02f2_$tag.json 2> gpurun_out/r02f2_$tag.err; }
run default A=1
run noskip TRAJOPT_B200_LS_SKIP=0
run ls2552 TRAJOPT_B200_LS=2,5,5,2
run ls2352 TRAJOPT_B200_LS=2,3,5,2
run ls2353 TRAJOPT_B200_LS=2,3,5,3
run ls2262 TRAJOPT_B200_LS=2,2,6,2
run ls3352 TRAJOPT_B200_LS=3,3,5,2
for w in forest bridge circle64 cross8; do
timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu > gpurun_out/r02f2_$w.json 2> gpurun_out/r02f2_$w.err
TRAJOPT_B200_LS_SKIP=0 timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu > gpurun_out/r02f2_${w}_noskip.json 2> gpurun_out/r02f2_${w}_noskip.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02f2_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02f2_")[1][:-5].ljust(16), "ms/step %.3f"%j["ms_per_step"], "e2e %.0f"%j["e2e"]["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_row_energy","k_row_grad","k_robot_ls","k_slack","k_bp_count","k_bp_fill") if n in k), "evals %.1fM"%(j["pairs_per_step"]["energy_plane_evals"]/1e6), "hist", [round(x,1) for x in j.get("ls_rung_hist_per_step",[])])
    except Exception as e:
        print(f, "ERR", e)
PY
