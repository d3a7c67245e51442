import csv, re, collections, sys
rows=csv.reader(open(sys.argv[1]))
corepath='/root/repo/pympc_quadruped_b200/csrc/mpcq_core.cuh'
cur_file=None; hdr=None; data=collections.defaultdict(lambda: collections.defaultdict(lambda:[0,0])); nfunc=0
stalls=collections.Counter()
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1]; continue
    if r[0]=='Function Name': nfunc = 1 if sys.argv[2] in r[1] else 0; continue
    if r[0]=='Line No': hdr=r; continue
    if hdr and cur_file and nfunc:
        try: ln=int(r[0])
        except: continue
        d=dict(zip(hdr,r))
        def _i(v):
            try: return int(v)
            except: return 0
        ie=_i(d.get('Instructions Executed','0')); sm=_i(d.get('# Samples','0'))
        data[cur_file][ln][0]+=ie; data[cur_file][ln][1]+=sm
for f in data: print(f, sum(v[0] for v in data[f].values()), sum(v[1] for v in data[f].values()))
core=[f for f in data if f.endswith('mpcq_core.cuh')][0]
src=open(corepath).read().split('\n')
funcs=[]
for i,l in enumerate(src,1):
    m=re.match(r'MPCQ_DEV [\w<>, ]*?\s*(\w+)\(',l)
    if m: funcs.append((i,m.group(1)))
funcs.append((len(src)+1,'end'))
tot_i=sum(v[0] for f in data for v in data[f].values()); tot_s=sum(v[1] for f in data for v in data[f].values())
print('total instr',tot_i,'samples',tot_s)
for (a,name),(b,_) in zip(funcs[:-1],funcs[1:]):
    ii=sum(v[0] for ln,v in data[core].items() if a<=ln<b); ss=sum(v[1] for ln,v in data[core].items() if a<=ln<b)
    if ii: print(f'{name:22s} lines {a}-{b-1}: instr {ii/tot_i*100:5.1f}%  samples {ss/tot_s*100:5.1f}%')
print('--- top lines by samples')
top=sorted(data[core].items(), key=lambda kv:-kv[1][1])[:45]
for ln,(ii,ss) in top: print(f'{ln:4d} instr {ii/tot_i*100:5.2f}% samp {ss/tot_s*100:5.2f}% | {src[ln-1].strip()[:110]}')
