"""Where does a kernel touch local memory?  From `nvdisasm -g -c kernel.cubin` (built with -lineinfo): every STL / LDL with its
loop nesting depth (backward branches define the loops) and source line, summed per (file, function-ish line).  Spills at
depth 0 are call-boundary / prologue saves (cheap); spills at depth >= 1 sit in loops.
   nvdisasm -g -c x.cubin > x.nvd;  python tools/sass_loop_spills.py x.nvd <kernel substring> [min depth]"""
import re
import sys
from collections import defaultdict


def main():
    path, kern = sys.argv[1], sys.argv[2]
    mind = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    inside, cur = False, None
    ins, labels = [], {}
    for ln in open(path, errors="replace"):
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.match(r"(\.L_x_\d+):", ln)
        if m:
            labels[m.group(1)] = len(ins)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            ins.append((cur, m.group(2).strip()))
    depth = [0] * len(ins)
    for i, (w, s) in enumerate(ins):
        m = re.search(r"\bBRA\b.*`\((\.L_x_\d+)\)", s)
        if m and m.group(1) in labels and labels[m.group(1)] <= i:
            for k in range(labels[m.group(1)], i + 1):
                depth[k] += 1
    per = defaultdict(lambda: [0, 0])
    tot = defaultdict(lambda: [0, 0])
    for i, (w, s) in enumerate(ins):
        op = 0 if re.search(r"\bSTL", s) else 1 if re.search(r"\bLDL", s) else None
        if op is None:
            continue
        tot[depth[i]][op] += 1
        if depth[i] >= mind:
            per[(w, depth[i])][op] += 1
    print("instructions", len(ins), " local-memory by loop depth (STL, LDL):", {d: tuple(v) for d, v in sorted(tot.items())})
    for (w, d), (a, b) in sorted(per.items(), key=lambda kv: (str(kv[0][0]), kv[0][1])):
        print(f"  {str(w):36s} depth {d}  STL {a:3d}  LDL {b:3d}")


if __name__ == "__main__":
    main()
