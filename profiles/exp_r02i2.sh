#!/bin/bash
# k_row_grad with a plane shared by two threads (128 registers, 16 warps per SM), new default line-search schedules: whole suite + all workloads
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02i2_tests.log 2>&1
cat gpurun_out/r02i2_tests.log
timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02i2_shard.json 2> gpurun_out/r02i2_shard.err
for w in forest bridge circle64 circle64c cross8; do
timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu > gpurun_out/r02i2_$w.json 2> gpurun_out/r02i2_$w.err
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02i2_batch.json 2> gpurun_out/r02i2_batch.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02i2_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02i2_")[1][:-5].ljust(10), "ms/step %.3f"%j["ms_per_step"], "value %.0f"%j["value"], "e2e %.0f"%j["e2e"]["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_row_energy","k_row_grad","k_robot_ls","k_slack","k_bp_count","k_bp_fill","k_bp_ccd","k_pack","k_piece","k_solve_bcr") if n in k))
    except Exception as e:
        print(f, "ERR", e)
PY
