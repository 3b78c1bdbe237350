"""Attribute an `ncu --page source --csv` dump (SASS view: one row per instruction) to SOURCE LINES without the .ncu-rep:
the line table comes from `nvdisasm -g -c` of the same cubin (built with -lineinfo).

  cuobjdump -xelf all libbpx.so && nvdisasm -g -c bpx_api.sm_100a.cubin > all.sass
  python tools/ncu_source_lines.py all.sass <mangled kernel name> gpurun_out/x.source.csv [top]
"""
import csv
import re
import sys
from collections import defaultdict


def line_table(sass_path, kernel):
    table, cur, inside = {}, None, False
    for ln in open(sass_path, errors="replace"):
        if ln.startswith(".text."):
            inside = ln.strip().rstrip(":") == ".text." + kernel
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    sass, kernel, src = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(sass, kernel)
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(rows[2][ia], 16)
    per_line = defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for r in rows[2:]:
        if len(r) <= max(ii, isamp):
            continue
        off = int(r[ia], 16) - base
        where = table.get(off, (None, ""))[0]
        n, s = int(r[ii] or 0), int(r[isamp] or 0)
        per_line[where][0] += n
        per_line[where][1] += s
        tot_i += n
        tot_s += s
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    for where, (n, s) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{str(where):40s} inst {n:12d} ({100.0 * n / max(tot_i, 1):5.1f} %)  samples {s:8d} ({100.0 * s / max(tot_s, 1):5.1f} %)")


if __name__ == "__main__":
    main()
