"""Like ncu_source_lines.py, aggregated by enclosing function of one source file (and top lines).
  python tools/ncu_by_function.py all.sass <kernel> x.source.csv <source file> [extra device functions...]"""
import bisect
import csv
import re
import sys
from collections import defaultdict

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from ncu_source_lines import line_table  # noqa: E402


def main():
    sass, kernel, srccsv, srcfile = sys.argv[1:5]
    tab = line_table(sass, kernel)
    rows = list(csv.reader(open(srccsv)))
    hdr = rows[1]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(rows[2][ia], 16)
    src = open(srcfile).read().split("\n")
    name = srcfile.split("/")[-1]
    starts = []
    for i, l in enumerate(src, 1):
        m = re.match(r"__host__ __device__ (?:__noinline__ |__forceinline__ |inline )*[\w:<>\*& ]+? (\w+)\(", l)
        if m:
            starts.append((i, m.group(1)))
    agg, lines = defaultdict(lambda: [0, 0]), defaultdict(lambda: [0, 0])
    for r in rows[2:]:
        if len(r) <= max(ii, isamp):
            continue
        w = tab.get(int(r[ia], 16) - base, (None, ""))[0]
        n, s = int(r[ii] or 0), int(r[isamp] or 0)
        if w is None:
            key = "?"
        elif w[0] != name:
            key = w[0]
        else:
            k = bisect.bisect_right([x[0] for x in starts], w[1]) - 1
            key = starts[k][1] if k >= 0 else "top"
        agg[key][0] += n
        agg[key][1] += s
        lines[w][0] += n
        lines[w][1] += s
    ti = sum(v[0] for v in agg.values())
    ts = sum(v[1] for v in agg.values())
    print("warp instructions", ti, "samples", ts)
    for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"{k:28s} inst {100 * n / max(ti, 1):5.1f}%  samples {100 * s / max(ts, 1):5.1f}%")
    for k, (n, s) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:14]:
        txt = src[k[1] - 1].strip()[:100] if k and k[0] == name else ""
        print(k, n, s, txt)


if __name__ == "__main__":
    main()
