#!/usr/bin/env python
"""Summarise an ncu report (raw page) into the key = value text kept under profiles/.  Usage: ncu_summary.py rep.ncu-rep out.txt "comment" """
import csv, subprocess, sys, io
rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = """dram__bytes_read.sum dram__bytes_write.sum gpu__time_duration.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed launch__block_size launch__grid_size launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__registers_per_thread launch__shared_mem_per_block_dynamic launch__waves_per_multiprocessor sm__cycles_elapsed.avg
sm__cycles_elapsed.avg.per_second sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__warps_active.avg.per_cycle_active smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__thread_inst_executed_per_inst_executed.ratio smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed
smsp__inst_executed_op_local_ld.sum smsp__inst_executed_op_local_st.sum""".split()
keys += sorted(k for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"))
lines = ["# " + note]
for k in keys:
    if k in d: lines.append("%s [%s] = %s" % (k, d[k][0], d[k][1]))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
