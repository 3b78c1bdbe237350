"""Split one kernel's cuobjdump -sass listing at RET / EXIT (the __noinline__ device functions of the kernel) and count, per
region, instructions / DFMA / local-memory traffic / MUFU / barriers: a quick look at where a kernel spills before going to
the GPU.   python tools/sass_regions.py kernel.sass"""
import re
import sys


def main():
    ins = []
    for l in open(sys.argv[1]):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append(m.group(2))
    regions, start = [], 0
    for i, s in enumerate(ins):
        if re.match(r"(@!?U?P\d\s+)?(RET|EXIT)", s.strip()) and not s.strip().startswith("@"):
            regions.append((start, i + 1))
            start = i + 1
    print(f"{'range':>14s} {'n':>6s} {'DFMA':>5s} {'DMMA':>5s} {'STL':>4s} {'LDL':>4s} {'MUFU':>4s} {'BAR':>4s} {'LDS':>5s} {'STS':>5s} {'LDG':>4s} {'STG':>4s} {'SHFL':>4s}")
    for a, b in regions:
        seg = ins[a:b]
        c = lambda p: sum(1 for s in seg if re.search(p, s))
        if b - a < 40:
            continue
        print(f"{a:6d}-{b:6d} {b - a:6d} {c(r'DFMA|DMUL|DADD'):5d} {c('DMMA'):5d} {c(r'STL'):4d} {c(r'LDL'):4d} {c('MUFU'):4d} {c(r'BAR'):4d} {c(r'LDS'):5d} {c(r'STS'):5d} {c(r'LDG'):4d} {c(r'STG'):4d} {c('SHFL'):4d}")


if __name__ == "__main__":
    main()
