#!/usr/bin/env python
"""Aggregate ncu_by_line.py output of skr_group_kernel by kernel phase (source line ranges found by marker comments)."""
import sys
lines_txt, src_path = sys.argv[1:3]
src = open(src_path).read().split('\n')
def find(pat):
    for i, l in enumerate(src):
        if pat in l:
            return i + 1
    raise KeyError(pat)
marks = [('fetch_unit', 'auto fetch_unit'), ('resolve/move (ctrl)', 'auto resolve_prev'), ('loop head', 'if (ctrl) fetch_unit(0)'), ('ctrl branch', 'if (ctrl) {'),
         ('worker init/zero', '// ==============='), ('load/expansion', 'if (!too_big) {'), ('toobig', 'if (too_big) {  // give up'),
         ('hash group', '// ---- group: claim'), ('leaders', '// ---- leaders and survivors'), ('publish', 'const uint32_t S = s_nsurv'),
         ('rank', '// ---- survivors ascending'), ('bitonic', 'uint32_t n2 = 1;'), ('offsets', '// ---- id offsets of the survivors'),
         ('mapv', '// ---- every instance learns'), ('mat setup/zero', 'constexpr int ITERS'), ('mat fill', 'uint32_t within[ITERS];'),
         ('suffix', '// suffix sums over the chunks of list'), ('ordered stage ids', 'out.stg_ids[idb + off[sI] + mat'), ('general', '// ---- id lists, general path'),
         ('stage table', "// ---- stage the unit's slice"), ('end', '// ------------------------------------------------------------------ bucket directory')]
marks = [(n, find(p)) for n, p in marks]
agg = {}
for l in open(lines_txt).read().splitlines()[3:]:
    p = l.split()
    loc, i, s = p[0], float(p[1]), float(p[2])
    f, ln = loc.rsplit(':', 1)
    ln = int(ln)
    name = f
    if f.startswith('skr_group.cu'):
        name = 'other(%d)' % ln
        for (n, a), (_, b) in zip(marks, marks[1:]):
            if a <= ln < b:
                name = n
        if ln < marks[0][1]:
            name = 'helpers(%d-)' % (ln // 50 * 50)
    agg.setdefault(name, [0, 0])
    agg[name][0] += i
    agg[name][1] += s
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[0] > 0.4 or v[1] > 0.4:
        print(f"{k:28s} inst {v[0]:6.2f}%  samples {v[1]:6.2f}%")
