"""Summarise an .ncu-rep (raw + source pages) into the few numbers the design notes use.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = page(rep, "raw")
    hdr, units, r = rows[0], rows[1], rows[2 + which]
    print("kernel:", r[hdr.index("Kernel Name")][:100])
    keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_active.avg"]
    for k in keys:
        if k in hdr:
            print("%-82s %-16s %s" % (k, units[hdr.index(k)], r[hdr.index(k)]))
    for i, h in enumerate(hdr):
        if "warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                if float(r[i]) > 0.2:
                    print("%-82s %s" % (h, r[i]))
            except ValueError:
                pass
    rows = page(rep, "source", ["--print-source", "sass"])
    # several kernels may follow each other; take the block of the requested kernel
    blocks, cur = [], None
    for row in rows:
        if row and row[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None:
            cur.append(row)
    blk = blocks[min(which, len(blocks) - 1)]
    hdr = blk[0]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [x for x in blk[1:] if len(x) == len(hdr)]

    def f(row, k):
        try:
            return float(row[ix[k]])
        except (ValueError, KeyError):
            return 0.0
    hist, tot = collections.Counter(), 0.0
    samp, ts = collections.Counter(), 0.0
    for row in data:
        src = row[ix["Source"]].strip()
        op = src.split()[1 if src.startswith("@") else 0] if src else "?"
        samp[op] += f(row, "# Samples")
        ts += f(row, "# Samples")
        if "LDS" in op or "STS" in op:
            ex, wf = f(row, "Instructions Executed"), f(row, "L1 Wavefronts Shared")
            if ex > 0:
                hist[(op, round(wf / ex, 1), round(f(row, "L1 Wavefronts Shared Ideal") / ex, 1))] += wf
                tot += wf
    print("shared-memory wavefronts by (op, wavefronts per instruction, ideal):")
    for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:10]:
        print("   %-28s %5.1f%%" % (k, 100 * v / tot))
    print("stall samples by opcode:", {k: "%.1f%%" % (100 * v / ts) for k, v in samp.most_common(8)})


if __name__ == "__main__":
    main()
