#!/bin/bash
# warp-level rings in k_narrow: tests, then A/B on the shard one of eight GPUs gets, forest, bridge
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_policy.py tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_edges.py tests/test_gpu_optplane.py -x -q 2>&1 | tail -15 ) > gpurun_out/r02w_tests.log
cat gpurun_out/r02w_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02w_$tag.json 2> gpurun_out/r02w_$tag.err; }
run default A=1
run nofilter TRAJOPT_B200_NP_FILTER=0
run occ5 TRAJOPT_B200_NP_OCC=5
run occ6 TRAJOPT_B200_NP_OCC=6
timeout 600 python bench.py --workload forest --steps 30 --warmup 5 --no-cpu > gpurun_out/r02w_forest.json 2> gpurun_out/r02w_forest.err
timeout 600 python bench.py --workload bridge --steps 30 --warmup 5 --no-cpu > gpurun_out/r02w_bridge.json 2> gpurun_out/r02w_bridge.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02w_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02w_")[1][:-5].ljust(10), "ms/step %.3f"%j["ms_per_step"], "e2e %.0f"%j["e2e"]["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_row_energy","k_row_grad","k_bp_ccd","k_bp_count","k_bp_fill","k_pack","k_robot_ls") if n in k))
    except Exception as e:
        print(f, "ERR", e)
PY
