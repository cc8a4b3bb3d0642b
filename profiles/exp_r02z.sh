#!/bin/bash
# source-level stall tables of the broadphase and pair kernels on the shard one of eight GPUs gets (one launch each)
mkdir -p gpurun_out
export TRAJOPT_B200_NO_GRAPH=1
for k in k_bp_count k_bp_fill k_narrow k_row_energy k_row_grad k_bp_ccd k_piece; do
  timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 3 -c 1 -k regex:^$k -o /tmp/z_$k -f python bench.py --emulate-rank 0/8 --steps 2 --warmup 3 --no-cpu > gpurun_out/r02z_ncu_$k.log 2>&1
  python profiles/hotlines.py /tmp/z_$k.ncu-rep $k 40 > gpurun_out/r02z_hot_$k.txt 2>&1
  ncu -i /tmp/z_$k.ncu-rep --page raw --csv > /tmp/z_$k.csv 2>/dev/null
  python - $k <<'PY' > gpurun_out/r02z_raw_$k.txt 2>&1
import csv,sys
k=sys.argv[1]
rows=list(csv.reader(open(f"/tmp/z_{k}.csv")))
hdr=rows[0]; vals=rows[2] if len(rows)>2 else rows[1]
want=["gpu__time_duration.sum","launch__registers_per_thread","launch__grid_size","launch__block_size","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","smsp__thread_inst_executed_per_inst_executed.ratio","dram__bytes_read.sum","dram__bytes_write.sum","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","sm__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum","l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","sm__inst_executed_pipe_lsu.sum"]
for h,v in zip(hdr,vals):
    if h in want or h.startswith("smsp__average_warp") and "per_issue" in h or "issue_stalled" in h and "per_warp_active" in h:
        print(h, v)
PY
done
ls -la gpurun_out | tail -25
