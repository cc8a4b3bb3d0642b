#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_policy.py -x -q 2>&1 | tail -15 ) > gpurun_out/r02t_tests.log
cat gpurun_out/r02t_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --emulate-rank 0/8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02t_$tag.json 2> gpurun_out/r02t_$tag.err; }
run default A=1
run pack4 TRAJOPT_B200_PACK_GRID=4
run pack8 TRAJOPT_B200_PACK_GRID=8
run pack32 TRAJOPT_B200_PACK_GRID=32
run nofilter TRAJOPT_B200_NP_FILTER=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02t_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k=j["kernels"]
        print(f.split("r02t_")[1][:-5].ljust(10), "ms/step %.3f"%j["ms_per_step"], "e2e %.0f"%j["e2e"]["value"], " ".join("%s=%.3f"%(n.replace("k_",""),k[n]["ms_per_step"]) for n in ("k_narrow","k_row_energy","k_row_grad","k_bp_ccd","k_bp_count","k_bp_fill","k_pack","k_robot_ls") if n in k))
    except Exception as e:
        print(f, "ERR", e)
PY
# SASS-level stall samples of k_narrow on the shard (one launch, after the warm-up iterations)
export TRAJOPT_B200_NO_GRAPH=1
ncu --set full --clock-control none --import-source on --launch-skip 3 -c 1 -k regex:k_narrow -o /tmp/np_shard -f python bench.py --emulate-rank 0/8 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02t_ncu_np.log 2>&1
python profiles/hotsass.py /tmp/np_shard.ncu-rep k_narrow 120 > gpurun_out/r02t_hotsass_k_narrow.txt 2>&1
python profiles/hotlines.py /tmp/np_shard.ncu-rep k_narrow 60 > gpurun_out/r02t_hotlines_k_narrow.txt 2>&1
ls -la gpurun_out | tail -5
