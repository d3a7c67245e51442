"""Static SASS instruction count per source function of one kernel:
   nvdisasm -g -c <cubin> > x.sass ; python tools/sass_static.py x.sass <kernel substring>"""
import re, sys, collections
core = '/root/repo/pympc_quadruped_b200/csrc/mpcq_core.cuh'
src = open(core).read().split('\n')
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r'MPCQ_DEV [\w<>, ]*?\s*(\w+)\(', l)
    if m: funcs.append((i, m.group(1)))
def fn_of(line):
    name = '?'
    for a, n in funcs:
        if a <= line: name = n
        else: break
    return name
cnt = collections.Counter(); inside = False; cur = ('?', 0); total = 0
for l in open(sys.argv[1]):
    if l.startswith('.text.'):
        inside = sys.argv[2] in l; continue
    if not inside: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        f, ln = m.group(1), int(m.group(2))
        cur = (fn_of(ln) if f.endswith('mpcq_core.cuh') else f.split('/')[-1], ln); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        cnt[cur[0]] += 1; total += 1
print('total', total)
for k, v in cnt.most_common(40): print(f'{k:28s} {v:6d}  {100*v/total:5.1f}%')
