#!/bin/bash
# k_piece with 128-thread CTAs where many blocks are in flight (same results): tests, shard, whole batch
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_policy.py tests/test_gpu_parity.py tests/test_gpu_batch.py -x -q 2>&1 | tail -5 ) > gpurun_out/r02o2_tests.log 2>&1
cat gpurun_out/r02o2_tests.log
TRAJOPT_B200_PIECE_CTA=128 timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02o2_shard_p128.json 2> gpurun_out/r02o2_shard_p128.err
TRAJOPT_B200_PIECE_CTA=384 timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02o2_shard_p384.json 2> gpurun_out/r02o2_shard_p384.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02o2_batch.json 2> gpurun_out/r02o2_batch.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02o2_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02o2_")[1][:-5].ljust(16), "ms/step %.3f"%j["ms_per_step"], "value %.0f"%j["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_piece","k_row_grad","k_row_energy") if n in k))
    except Exception as e:
        print(f, "ERR", e)
PY
