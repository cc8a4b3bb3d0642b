#!/bin/bash
# chunks of 1024 candidates in k_narrow
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_policy.py tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_edges.py tests/test_gpu_optplane.py -x -q 2>&1 | tail -5 ) > gpurun_out/r02m2_tests.log 2>&1
cat gpurun_out/r02m2_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02m2_$tag.json 2> gpurun_out/r02m2_$tag.err; }
run default A=1
run chunk512 TRAJOPT_B200_NP_CHUNK=512
run pmem6 TRAJOPT_B200_NP_PMEM=6
run pmem4 TRAJOPT_B200_NP_PMEM=4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02m2_batch.json 2> gpurun_out/r02m2_batch.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02m2_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02m2_")[1][:-5].ljust(16), "ms/step %.3f"%j["ms_per_step"], "value %.0f"%j["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_pack","k_bp_top+k_np_top") if n in k))
    except Exception as e:
        print(f, "ERR", e)
PY
