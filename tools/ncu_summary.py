#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box): key metrics per profiled launch, SASS opcode mix and the
instructions with the most stall samples.  Usage: tools/ncu_summary.py report.ncu-rep [launch_index] > profiles/x.txt"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = re.compile(
    r"^(gpu__time_duration\.sum|smsp__inst_executed\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"launch__registers_per_thread|launch__grid_size|launch__block_size|launch__shared_mem_per_block_dynamic|launch__waves_per_multiprocessor|"
    r"launch__occupancy_limit_(registers|shared_mem|warps)|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"smsp__warps_eligible\.avg\.per_cycle_active|sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__cycles_elapsed\.avg|lts__t_sector_hit_rate\.pct|l1tex__t_sectors_pipe_lsu_mem_global_op_(ld|st)\.sum|"
    r"l1tex__t_requests_pipe_lsu_mem_global_op_(ld|st)\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
    r"smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio|sm__inst_executed_pipe_[a-z0-9_]+\.sum)$")


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    names = hdr.index("Kernel Name")
    for li, row in enumerate(raw[2:]):
        print(f"=== launch {li}: {row[names]}")
        stalls = []
        for i, h in enumerate(hdr):
            if KEYS.match(h):
                if "issue_stalled" in h:
                    try:
                        stalls.append((float(row[i]), h))
                    except ValueError:
                        pass
                else:
                    print(f"  {h:75s} {row[i]:>18s} {units[i]}")
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f"  {h:75s} {v:18.3f}")
    src = run(["-i", rep, "--page", "source", "--csv"])
    # the source page concatenates one table per launch, each preceded by a "Kernel Name" line
    blocks = re.split(r'(?m)^"Kernel Name",', src)
    for bi, blk in enumerate(blocks[1:]):
        lines = list(csv.reader(io.StringIO(blk)))
        kname = lines[0][0]
        h = lines[1]
        data = [r for r in lines[2:] if len(r) == len(h)]
        isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        tot = sum(int(r[iex]) for r in data)
        samp = sum(int(r[isamp]) for r in data)
        print(f"\n=== source page, launch {bi}: {kname}\n  static SASS instructions {len(data)}, executed warp-instructions {tot}, stall samples {samp}")
        ops = collections.Counter()
        for r in data:
            p = r[isrc].split()
            op = (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
            ops[op] += int(r[iex])
        print("  opcode mix (share of executed warp-instructions):")
        print("   " + "  ".join(f"{o} {100*c/tot:.1f}%" for o, c in ops.most_common(18)))
        print("  top stall-sample instructions:")
        for r in sorted(data, key=lambda r: -int(r[isamp]))[:14]:
            print(f"   {int(r[isamp]):7d} ({100*int(r[isamp])/max(samp,1):4.1f}%)  exec {int(r[iex]):10d}  {r[isrc].strip()[:80]}")


if __name__ == "__main__":
    main()
