#!/bin/bash
# accepted-rung histogram of the line search and policies with a short round 1
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_policy.py tests/test_gpu_parity.py -x -q 2>&1 | tail -8 ) > gpurun_out/r02d2_tests.log
cat gpurun_out/r02d2_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02d2_$tag.json 2> gpurun_out/r02d2_$tag.err; }
run default A=1
run ls2552 TRAJOPT_B200_LS=2,5,5,2
run ls2553 TRAJOPT_B200_LS=2,5,5,3
run ls2562 TRAJOPT_B200_LS=2,5,6,2
run ls2362 TRAJOPT_B200_LS=2,3,6,2
run ls2352 TRAJOPT_B200_LS=2,3,5,2
run ls3552 TRAJOPT_B200_LS=3,5,5,2
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02d2_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02d2_")[1][:-5].ljust(10), "ms/step %.3f"%j["ms_per_step"], "e2e %.0f"%j["e2e"]["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_row_energy","k_row_grad","k_robot_ls","k_slack") if n in k), "evals %.1fM"%(j["pairs_per_step"]["energy_plane_evals"]/1e6), "hist", [round(x,1) for x in j.get("ls_rung_hist_per_step",[])])
    except Exception as e:
        print(f, "ERR", e)
PY
