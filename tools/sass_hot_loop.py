#!/usr/bin/env python
"""Instruction mix of the hot loop of sw_wavefront_kernel<16,3,1> in a built library (no GPU needed).

    python tools/sass_hot_loop.py nanomonsv_b200/libnmsv_sw.so [mangled kernel name] [v]

Finds the backward branch that encloses the most VIADDMNMX instructions (the unrolled step body), drops the regions a
forward predicated branch jumps over (the exact path of the running maximum) and prints the per-iteration counts:
`fast` = instructions on the common path, `dpx` = VIADDMNMX / VIMNMX*, `other` = the rest. With a third argument the
non-DPX instructions are listed. Divide by the unroll factor (8) for per-step numbers.
"""
import re,collections,subprocess,sys
lib=sys.argv[1]; fn=sys.argv[2] if len(sys.argv)>2 else '_ZN4nmsv19sw_wavefront_kernelILi16ELi3ELi1EEEvNS_8SwParamsE'
out=subprocess.run(['cuobjdump','-sass','-fun',fn,lib],capture_output=True,text=True).stdout
ops=[]
for l in out.splitlines():
    m=re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);",l)
    if m: ops.append((int(m.group(1),16),m.group(3),m.group(4).strip(),m.group(2) or ''))
best=None
for i,(a,o,r,pr) in enumerate(ops):
    if o.startswith('BRA'):
        m=re.search(r"0x([0-9a-f]+)",r)
        if m:
            t=int(m.group(1),16)
            if t<a:
                body=[x for x in ops if t<=x[0]<=a]
                n=sum(1 for x in body if x[1].startswith('VIADDMNMX'))
                # the innermost loop that holds the unrolled step body (8 steps x 66 DPX = 528 VIADDMNMX/VIMNMX3 of which
                # 400 are VIADDMNMX): the smallest body among the loops with at least 300 of them
                if n>=300 and (best is None or len(body)<len(best[3])): best=(n,t,a,body)
n,t,a,body=best
# remove forward-skipped regions: @!P BRA target forward inside loop => exact path skipped
fast=[];skip_until=None
for x in body:
    if skip_until is not None:
        if x[0]<skip_until: continue
        skip_until=None
    fast.append(x)
    if x[1]=='BRA' and x[3].strip().startswith('@'):
        m=re.search(r"0x([0-9a-f]+)",x[2]); tg=int(m.group(1),16)
        if tg>x[0] and tg<=a: skip_until=tg
c=collections.Counter(x[1] for x in fast)
dpx=sum(v for k,v in c.items() if k.startswith(('VIADDMNMX','VIMNMX')))
print("loop %#x-%#x total %d fast %d dpx %d other %d"%(t,a,len(body),len(fast),dpx,len(fast)-dpx))
print(c.most_common())
if len(sys.argv)>3:
    for x in fast:
        if not x[1].startswith(('VIADDMNMX','VIMNMX3')): print(hex(x[0]),x[3],x[1],x[2])
