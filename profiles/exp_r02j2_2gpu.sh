#!/bin/bash
# two GPUs: the multi-GPU tests (sharded == single bitwise: decoupled, coupled, unequal shares, forced overflow; C++ host on two GPUs) and bench lines
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_host_dropin.py -q -k "sharded or two_gpus" 2>&1 | tail -12 ) > gpurun_out/r02j2_tests_2gpu.log 2>&1
cat gpurun_out/r02j2_tests_2gpu.log
for w in batch circle64 circle64c; do
  a=""; [ $w != batch ] && a="--workload $w"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 $a --no-cpu > gpurun_out/r02j2_bench_${w}_2gpu.json 2> gpurun_out/r02j2_bench_${w}_2gpu.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02j2_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r02j2_bench_")[1][:-5].ljust(16), j["n_gpus"], "ms/step %.3f"%j["ms_per_step"], "value %.0f"%j["value"], "e2e %.0f"%j["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e)
PY
