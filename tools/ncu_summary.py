"""Summarise one kernel of an .ncu-rep (ncu --set full ... -o file) as the small CSV kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "command that was profiled" [--index K] > profiles/rNN_ncu_full_<kernel>.csv

--index K picks the K-th profiled launch of the report (default 0).

Prints `metric,unit,value` rows for the launch configuration, time, DRAM traffic, pipe utilisation and issue
statistics, then the instruction mix (top opcodes, share of executed warp instructions) and the warp-stall mix
from the source page.  Reads the report with `ncu -i ... --page raw|source --csv`; no GPU needed.
"""
import collections
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def page(rep, name):
    return list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True,
                                          text=True, check=True).stdout.splitlines()))


def main():
    rep = sys.argv[1]
    what = sys.argv[2] if len(sys.argv) > 2 else ""
    index = int(sys.argv[sys.argv.index("--index") + 1]) if "--index" in sys.argv else 0
    rows = page(rep, "raw")
    hdr, units, val = rows[0], rows[1], rows[2 + index]
    print("# ncu --set full --clock-control none, one launch: %s" % what)
    print("metric,unit,value")
    print('Kernel Name,,"%s"' % val[hdr.index("Kernel Name")])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print('%s,%s,"%s"' % (w, units[i], val[i]))
    src = page(rep, "source")
    h = None
    sect = -1
    byop, stalls, total = collections.Counter(), collections.Counter(), 0
    for r in src:
        if len(r) > 3 and r[0] == "Address":
            sect += 1                                      # two "Address, Source, ..." sections per profiled launch
            if sect > 2 * index:
                break
            if sect < 2 * index:
                h = None
                continue
            h = r
            i_src, i_exec = h.index("Source"), h.index("Instructions Executed")
            cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
            continue
        if h and len(r) == len(h):
            try:
                e = int(r[i_exec])
            except ValueError:
                continue
            tok = r[i_src].strip().split()
            op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
            byop[op] += e
            total += e
            for i in cols:
                try:
                    stalls[h[i]] += int(r[i])
                except ValueError:
                    pass
    print('instruction_mix,%%,"%s"' % " ".join("%s:%.1f" % (o, 100.0 * c / total) for o, c in byop.most_common(14)))
    st = sum(stalls.values())
    print('warp_stall_mix,%%,"%s"' % " ".join("%s:%.1f" % (k[6:], 100.0 * v / st) for k, v in stalls.most_common(9)))


if __name__ == "__main__":
    main()
