#!/bin/bash
# usage: sass_spills_by_function.sh x.nvd <kernel substring> [min depth]   -- sass_loop_spills.py aggregated by source function
python /root/repo/tools/sass_loop_spills.py "$1" "$2" "${3:-1}" | python -c "
import sys,re,collections,bisect
src=open('/root/repo/itensornetworksnext.jl_b200/csrc/bpx_apply3.cuh').read().split('\n')
starts=[]
for i,l in enumerate(src,1):
    m=re.match(r'__host__ __device__ (?:__noinline__ |__forceinline__ |BPX_JACOBI_INLINE |inline )*[\w:<>\*& ]+? (\w+)\(',l)
    if m: starts.append((i,m.group(1)))
agg=collections.defaultdict(lambda:[0,0])
for l in sys.stdin:
    m=re.match(r\"\s+\('(\S+)', (\d+)\)\s+depth (\d+)\s+STL\s+(\d+)\s+LDL\s+(\d+)\",l)
    if not m:
        if l.startswith('instructions'): print(l.strip())
        continue
    f,ln,d,a,b=m.group(1),int(m.group(2)),int(m.group(3)),int(m.group(4)),int(m.group(5))
    if f=='bpx_apply3.cuh':
        k=bisect.bisect_right([x[0] for x in starts],ln)-1
        key=starts[k][1]
    else: key=f
    agg[(key,d)][0]+=a; agg[(key,d)][1]+=b
for k,v in sorted(agg.items()): print(' ',k,v)
"
