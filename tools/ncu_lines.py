"""Joins the SASS-level source page of an ncu report with nvdisasm line info: per source line, executed warp
instructions, active lanes per instruction and stall samples.
usage: ncu_lines.py <report.ncu-rep> <object-or-so with the cubin> <kernel-name-substring> [top-N] [line-regex]"""
import re,csv,sys,subprocess
from collections import defaultdict
rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
subprocess.run(f"mkdir -p /tmp/cub2 && cd /tmp/cub2 && rm -f *.cubin dis.txt && cuobjdump -xelf all {obj} >/dev/null && for f in *.cubin; do nvdisasm -g -c $f >> dis.txt 2>/dev/null; done", shell=True, check=True)
subprocess.run(f"ncu -i {rep} --page source --csv > /tmp/cub2/src.csv 2>/dev/null", shell=True, check=True)
lines=open('/tmp/cub2/dis.txt').read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.') and kern in l][0]
end=[i for i,l in enumerate(lines) if l.startswith('//---') and i>start]
end=end[0] if end else len(lines)
cur=None; ins=[]
for l in lines[start:end]:
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2)), m.group(3)); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16), cur, m.group(2)))
rows=list(csv.reader(open('/tmp/cub2/src.csv')))
h=rows[1]; ii=h.index('Instructions Executed'); ti=h.index('Thread Instructions Executed'); si=h.index('# Samples')
data=rows[2:]
data=[r for r in data if len(r)>max(ii,ti,si) and r[ii].isdigit()][:len(ins)]   # several launches in the report: the first one
assert len(data)==len(ins),(len(data),len(ins))
agg=defaultdict(lambda:[0,0,0])
for r,(off,cur,txt) in zip(data,ins):
    a=agg[(cur[0],cur[1])] if cur else agg[('?',0)]
    a[0]+=int(r[ii]); a[1]+=int(r[ti]); a[2]+=int(r[si])
tot=sum(a[0] for a in agg.values()); tott=sum(a[1] for a in agg.values())
print('total warp inst',tot,'eff',tott/tot)
byfile=defaultdict(lambda:[0,0,0])
for (f,l),a in agg.items():
    for k in range(3): byfile[f][k]+=a[k]
for f,a in byfile.items(): print(f, a[0], f"{a[0]/tot*100:.1f}%", 'eff', round(a[1]/max(a[0],1),1), 'samples',a[2])
print('--- top lines')
N=int(sys.argv[4]) if len(sys.argv)>4 else 40
for (f,l),a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:N]:
    print(f"{f}:{l}".ljust(34), f"{a[0]/tot*100:5.1f}% eff {a[1]/max(a[0],1):5.1f} samp {a[2]} n={a[0]}")
if len(sys.argv)>5:
    pat=sys.argv[5]
    for r,(off,cur,txt) in zip(data,ins):
        if cur and re.search(pat, f"{cur[0]}:{cur[1]}"):
            ie=int(r[ii]); te=int(r[ti]); print(f"{off:6x} {cur[0][:22]}:{cur[1]:<4d} {ie:>10d} eff {te/max(ie,1):5.1f}  {txt[:70]}")
