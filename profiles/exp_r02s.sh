#!/bin/bash
# A/B of the kernel variants on the share one of eight GPUs gets of the 1024-problem batch
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_policy.py tests/test_gpu_parity.py tests/test_gpu_batch.py -x -q 2>&1 | tail -15 ) > gpurun_out/r02s_tests.log
cat gpurun_out/r02s_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02s_$tag.json 2> gpurun_out/r02s_$tag.err; }
run default A=1
run nofilter TRAJOPT_B200_NP_FILTER=0
run enocc0 TRAJOPT_B200_EN_OCC=0
run enocc2 TRAJOPT_B200_EN_OCC=2
run ccd4 TRAJOPT_B200_CCD_OCC=4
run ccd12 TRAJOPT_B200_CCD_OCC=12
run ls2216 TRAJOPT_B200_LS=2,2,16
run ls2210 TRAJOPT_B200_LS=2,2,10
run ls239 TRAJOPT_B200_LS=2,3,9
run ls246 TRAJOPT_B200_LS=2,4,6
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02s_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02s_")[1][:-5].ljust(10), "ms/step %.3f"%j["ms_per_step"], "e2e %.0f"%j["e2e"]["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_row_energy","k_row_grad","k_bp_ccd","k_bp_count","k_bp_fill","k_pack","k_robot_ls") if n in k), "exact", j["pairs_per_step"].get("np_kdop_exact"))
    except Exception as e:
        print(f, "ERR", e)
PY
