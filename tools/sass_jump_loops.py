#!/usr/bin/env python
"""Per-loop instruction counts of the sw_jump warp kernel's matrix functions (no GPU needed):
    python tools/sass_jump_loops.py nanomonsv_b200/libnmsv_sw.so
Lists every innermost backward-branch loop that holds SHF.L.W (path-bit) instructions: two steps of S cells each."""
import re, collections, subprocess, sys
lib = sys.argv[1]
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
fn = None; funcs = collections.OrderedDict()
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: fn = m.group(1); funcs[fn] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", l)
    if m and fn: funcs[fn].append((int(m.group(1), 16), m.group(3), m.group(4).strip()))
for fn, ops in funcs.items():
    if 'jump_matrix_any' not in fn and 'sw_jump_warp' not in fn: continue
    loops = []
    for a, o, r in ops:
        if o.startswith('BRA'):
            m = re.search(r"0x([0-9a-f]+)", r)
            if m and int(m.group(1), 16) < a: loops.append((int(m.group(1), 16), a))
    inner = [lp for lp in loops if not any(o != lp and lp[0] <= o[0] and o[1] <= lp[1] for o in loops)]
    print(fn, len(ops), 'instructions')
    for t, a in inner:
        body = [x for x in ops if t <= x[0] <= a]
        shf = sum(1 for x in body if x[1].startswith('SHF.L.W'))
        if shf < 4: continue
        c = collections.Counter(x[1].split('.')[0] for x in body)
        cells = shf // 2 if 'ILb0' in fn else shf // 3
        print("  loop %#x-%#x: %d instr, %d path shifts (~%d cells) -> %.1f instr/cell  %s" % (t, a, len(body), shf, cells, len(body) / max(cells, 1), dict(c.most_common(8))))
