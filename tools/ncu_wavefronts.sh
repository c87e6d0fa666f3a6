#!/bin/bash
# shared-memory wavefronts, instructions, FP64 / DMMA pipe use, stall mix and duration of the zip forward kernel launches of one zip_bench configuration
# usage: tools/ncu_wavefronts.sh <label> <workload> <sweep-spec>   (IMC_LIB_PATH selects an experiment build)
label=$1; wl=$2; spec=$3
ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.per_cycle_active,sm__cycles_active.avg,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio \
    --clock-control none -k regex:zip_forward_kernel -c 6 --csv --log-file /tmp/wf_$label.csv \
    python tools/zip_bench.py --workload $wl --check 0 --reps 1 --sweep $spec > /tmp/wf_$label.log 2>&1
python - "$label" <<'PY'
import csv, sys
label = sys.argv[1]
rows = [r for r in csv.reader(open('/tmp/wf_%s.csv' % label)) if len(r) > 5]
hdr = [r for r in rows if r[0] == 'ID'][0]
n, v, k = hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
num = lambda x: float(x.replace(',', ''))
dur = {r[k]: num(r[v]) for r in rows if r[0] != 'ID' and r[n] == 'gpu__time_duration.sum'}
top = max(dur, key=dur.get)          # (every call also launches the -- usually empty -- plain-form pass)
print(label, {r[n].replace('l1tex__data_pipe_lsu_wavefronts_mem_shared', 'smem_wf'): r[v] for r in rows if r[0] == top})
PY
