#!/bin/bash
# chunk size and occupancy of k_narrow under on-demand chunks
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02l2_$tag.json 2> gpurun_out/r02l2_$tag.err; }
run default A=1
run chunk256 TRAJOPT_B200_NP_CHUNK=256
run chunk128 TRAJOPT_B200_NP_CHUNK=128
run pmem4 TRAJOPT_B200_NP_PMEM=4
run pmem6 TRAJOPT_B200_NP_PMEM=6
run pmem6c256 TRAJOPT_B200_NP_PMEM=6 TRAJOPT_B200_NP_CHUNK=256
TRAJOPT_B200_NP_CHUNK=256 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02l2_batch_chunk256.json 2> gpurun_out/r02l2_batch_chunk256.err
TRAJOPT_B200_NP_CHUNK=128 timeout 600 python bench.py --workload forest --steps 30 --warmup 5 --no-cpu > gpurun_out/r02l2_forest_chunk128.json 2> gpurun_out/r02l2_forest_chunk128.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02l2_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02l2_")[1][:-5].ljust(16), "ms/step %.3f"%j["ms_per_step"], "value %.0f"%j["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_pack","k_bp_top+k_np_top") if n in k))
    except Exception as e:
        print(f, "ERR", e)
PY
