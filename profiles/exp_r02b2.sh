#!/bin/bash
# k_narrow: one order-preserving compaction per chunk + k_np_count; broadphase: ballot decode of the item leaf, two items in flight
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02b2_tests.log 2>&1
cat gpurun_out/r02b2_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02b2_$tag.json 2> gpurun_out/r02b2_$tag.err; }
run default A=1
timeout 600 python bench.py --workload forest --steps 30 --warmup 5 --no-cpu > gpurun_out/r02b2_forest.json 2> gpurun_out/r02b2_forest.err
timeout 600 python bench.py --workload bridge --steps 30 --warmup 5 --no-cpu > gpurun_out/r02b2_bridge.json 2> gpurun_out/r02b2_bridge.err
timeout 600 python bench.py --workload circle64 --steps 30 --warmup 5 --no-cpu > gpurun_out/r02b2_circle64.json 2> gpurun_out/r02b2_circle64.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02b2_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02b2_")[1][:-5].ljust(10), "ms/step %.3f"%j["ms_per_step"], "e2e %.0f"%j["e2e"]["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_row_energy","k_row_grad","k_bp_ccd","k_bp_count","k_bp_fill","k_pack","k_robot_ls","k_piece","k_solve_bcr","k_slack") if n in k))
    except Exception as e:
        print(f, "ERR", e)
PY
