import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
H=None; data=[]
for r in rows:
    if 'Instructions Executed' in r: H=r; continue
    if H and len(r)>10: data.append(r)
si=H.index('Source'); ie=H.index('Instructions Executed'); ss=H.index('# Samples'); te=H.index('Thread Instructions Executed')
tot=sum(int(r[ie]) for r in data); tots=sum(int(r[ss]) for r in data)
print("SASS instrs", len(data), "total warp-inst", tot)
seg=0; acc=0; accs=0; acct=0; start=0
for k,r in enumerate(data):
    acc+=int(r[ie]); accs+=int(r[ss]); acct+=int(r[te])
    if 'BAR.SYNC' in r[si] or k==len(data)-1:
        if acc/tot > 0.004:
            print(f"seg {seg:2d} sass[{start:5d}-{k:5d}] inst {acc/tot*100:5.1f}% ({acc/1e6:6.2f}M) samples {accs/tots*100:5.1f}% lanes/inst {acct/max(acc,1):5.1f}  barrier-exec {int(r[ie])}")
        seg+=1; acc=0; accs=0; acct=0; start=k+1
