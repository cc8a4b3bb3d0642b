#!/bin/bash
# line-search schedules after the bound mask (every search accepts its first evaluated rung on the batch)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_policy.py -x -q 2>&1 | tail -5 ) > gpurun_out/r02h2_tests.log 2>&1
cat gpurun_out/r02h2_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02h2_$tag.json 2> gpurun_out/r02h2_$tag.err; }
run default A=1
run ls2533 TRAJOPT_B200_LS=2,5,3,3
run ls2523 TRAJOPT_B200_LS=2,5,2,3
run ls2532 TRAJOPT_B200_LS=2,5,3,2
run ls2543 TRAJOPT_B200_LS=2,5,4,3
for w in forest bridge circle64; do
for pol in 9,9,2 2,9,2 3,9,2 5,9,2 2,9,3,3 3,9,3,5; do
TRAJOPT_B200_LS=$pol timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu > gpurun_out/r02h2_${w}_$pol.json 2> gpurun_out/r02h2_${w}_$pol.err
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02h2_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02h2_")[1][:-5].ljust(18), "ms/step %.3f"%j["ms_per_step"], "e2e %.0f"%j["e2e"]["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_row_energy","k_row_grad","k_robot_ls","k_slack") if n in k), "evals %.1fM"%(j["pairs_per_step"]["energy_plane_evals"]/1e6), "hist", [round(x,1) for x in j.get("ls_rung_hist_per_step",[])][:4])
    except Exception as e:
        print(f, "ERR", e)
PY
