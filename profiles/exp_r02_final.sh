#!/bin/bash
# end of round 2: whole GPU suite, the default bench line (with cpu_baseline) and the reference arm as the driver runs them,
# the other workloads, then the ncu captures (profiles/run_ncu_r02b.sh)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r02final_tests.log 2>&1
cat gpurun_out/r02final_tests.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02final_bench_reference.json 2> gpurun_out/r02final_bench_reference.err
timeout 900 python bench.py > gpurun_out/r02final_bench_default.json 2> gpurun_out/r02final_bench_default.err
for w in forest bridge cross8 circle64 circle64c; do
timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/r02final_bench_$w.json 2> gpurun_out/r02final_bench_$w.err
done
timeout 600 python bench.py --emulate-rank 0/8 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02final_bench_shard0of8.json 2> gpurun_out/r02final_bench_shard0of8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02final_bench_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        cb=j.get("cpu_baseline") or {}
        print(f.split("r02final_bench_")[1][:-5].ljust(12), j.get("impl","b200"), "ms/step %.3f"%j["ms_per_step"], "value %.0f"%j["value"], "e2e %.0f"%j["e2e"]["value"], "cpu %s (%s cores)"%(cb.get("value"), cb.get("cores")), "roof", (j.get("roofline") or {}).get("kernel"), (j.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
bash profiles/run_ncu_r02b.sh r02b
ls gpurun_out | grep r02b | head -40
