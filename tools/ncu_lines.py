#!/usr/bin/env python
"""Per-source-line view of an ncu capture (development aid): joins the SASS page of a .ncu-rep (instructions executed,
stall samples) with the line info `nvdisasm -g` prints for the same kernel in the in-tree library.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep 'k_tree2ILi10ELi3ELi2ELi0ELb1ELb1ELb1ELi1536' [--top 40]

The join is positional (k-th SASS instruction of the kernel in both listings), so the library must be the build
that was profiled.
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal-ghm_b200", "ghm_b200", "libghm_b200.so")


def disasm_lines(mangled_part):
    tmp = tempfile.mkdtemp(prefix="cub_")
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    for f in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
        elf = subprocess.run(["cuobjdump", "-elf", f], capture_output=True, text=True).stdout
        if mangled_part not in elf:
            continue
        txt = subprocess.run(["nvdisasm", "-g", f], capture_output=True, text=True).stdout
        out, on, cur = [], False, (None, 0)
        for ln in txt.splitlines():
            if ln.startswith(".text."):
                on = mangled_part in ln
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.append((cur, m.group(2).strip()))
        if out:
            return out
    raise SystemExit("kernel not found in " + LIB)


def main():
    rep, part = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    sass = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    dis = disasm_lines(part)
    if len(dis) != len(sass):
        sys.stderr.write("warning: %d SASS instructions in the report, %d in the library\n" % (len(sass), len(dis)))
    agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    tot_i = tot_s = 0
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    for k, r in enumerate(sass[:len(dis)]):
        (fn, line), op = dis[k]
        inst = int(r[col["Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
        a = agg[(fn, line)]
        a[0] += inst
        a[1] += smp
        a[2] += 1
        for n in stall_cols:
            v = int(r[col[n]] or 0)
            if v:
                a[3][n] += v
        tot_i += inst
        tot_s += smp
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    srcs = {}
    for (fn, line), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        if fn not in srcs:
            p = os.path.join(ROOT, "multimodal-ghm_b200", "csrc", fn or "")
            srcs[fn] = open(p).read().splitlines() if fn and os.path.exists(p) else []
        text = srcs[fn][line - 1].strip()[:70] if 0 < line <= len(srcs[fn]) else ""
        st = ", ".join("%s %.0f%%" % (n[6:], 100.0 * v / max(a[1], 1)) for n, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
        print("%5.1f%% smp %5.1f%% inst %4d sass  %s:%d  %s   [%s]" % (100.0 * a[1] / tot_s, 100.0 * a[0] / tot_i, a[2], fn, line, text, st))


if __name__ == "__main__":
    main()
