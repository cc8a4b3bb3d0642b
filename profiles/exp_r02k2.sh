#!/bin/bash
# k_narrow hands its chunks out on demand: tests, shard, whole batch, forest
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_policy.py tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_edges.py tests/test_gpu_optplane.py -x -q 2>&1 | tail -5 ) > gpurun_out/r02k2_tests.log 2>&1
cat gpurun_out/r02k2_tests.log
timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02k2_shard.json 2> gpurun_out/r02k2_shard.err
timeout 600 python bench.py --emulate-rank 3/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02k2_shard3.json 2> gpurun_out/r02k2_shard3.err
timeout 600 python bench.py --workload forest --steps 30 --warmup 5 --no-cpu > gpurun_out/r02k2_forest.json 2> gpurun_out/r02k2_forest.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02k2_batch.json 2> gpurun_out/r02k2_batch.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02k2_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02k2_")[1][:-5].ljust(10), "ms/step %.3f"%j["ms_per_step"], "value %.0f"%j["value"], "e2e %.0f"%j["e2e"]["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_row_energy","k_row_grad","k_bp_count","k_pack","k_piece") if n in k))
    except Exception as e:
        print(f, "ERR", e)
PY
