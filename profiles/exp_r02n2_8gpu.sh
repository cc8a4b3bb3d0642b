#!/bin/bash
# eight GPUs: sharded == single at 8 ranks (bitwise; decoupled, coupled, unequal shares, forced overflow) and the bench lines of the
# workloads that shard (1024-problem batch: strong scaling; 64 UAVs sharded, decoupled and coupled)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "sharded and 8" 2>&1 | tail -6 ) > gpurun_out/r02n2_tests_8gpu.log 2>&1
cat gpurun_out/r02n2_tests_8gpu.log
for n in 8 4; do
for w in batch circle64; do
  a=""; [ $w != batch ] && a="--workload $w"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 $a --no-cpu > gpurun_out/r02n2_bench_${w}_${n}gpu.json 2> gpurun_out/r02n2_bench_${w}_${n}gpu.err
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02n2_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r02n2_bench_")[1][:-5].ljust(16), j["n_gpus"], "ms/step %.3f"%j["ms_per_step"], "value %.0f"%j["value"], "e2e %.0f"%j["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e)
PY
