#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): headline metrics of every profiled launch and,
per code region of the hottest kernel, the share of executed warp instructions and stall samples.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--regions]
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    for r in data:
        print("==", r[name_i])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"   {m:80s} {r[i]:>16s} {units[i]}")
    if "--regions" in sys.argv:
        rows = ncu_csv(rep, "source")
        hdr = rows[1]
        data = [r for r in rows[2:] if len(r) > 8 and r[0].startswith("0x")]
        ia, isrc = hdr.index("Address"), hdr.index("Source")
        iex, ith, ismp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        first = int(data[0][ia], 16)
        seen, out = set(), []
        for r in data:
            a = int(r[ia], 16) - first
            if a in seen:
                break  # second launch of the same kernel
            seen.add(a)
            out.append((a, int(r[iex]), int(r[ith]), int(r[ismp]), r[isrc].strip()))
        tot = sum(o[1] for o in out)
        tots = sum(o[3] for o in out)
        regs, cur = [], [out[0]]
        for o in out[1:]:
            p = cur[-1][1]
            if (p == 0 and o[1] == 0) or (p > 0 and abs(o[1] - p) / p < 0.25):
                cur.append(o)
            else:
                regs.append(cur)
                cur = [o]
        regs.append(cur)
        print(f"total warp instructions {tot}, stall samples {tots}")
        for g in regs:
            ex, th, sm = sum(x[1] for x in g), sum(x[2] for x in g), sum(x[3] for x in g)
            if ex / tot > 0.004 or sm / max(tots, 1) > 0.01:
                print(f"  {g[0][0]:#07x}-{g[-1][0]:#07x} n={len(g):4d} exec/inst={g[0][1]:9d} inst={ex / tot * 100:5.1f}% "
                      f"lanes={th / max(ex, 1):5.1f} stall_samples={sm / max(tots, 1) * 100:5.1f}%  {g[0][4][:44]}")


if __name__ == "__main__":
    main()
