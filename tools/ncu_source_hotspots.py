import csv, sys, collections
path=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
rows=list(csv.reader(open(path)))
cur=None; data=[]
H=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Line No": H=r; continue
    if H and r[0] not in ("","Function Name") and r[0].isdigit():
        ie=H.index('Instructions Executed'); ss=H.index('# Samples')
        try: data.append((int(r[ie]), int(r[ss]), cur, int(r[0]), r[1].strip()))
        except Exception as e: pass
tot=sum(d[0] for d in data); tots=sum(d[1] for d in data)
print("total warp-inst", tot, "samples", tots)
for d in sorted(data, key=lambda x:-x[1])[:topn]:
    print(f"{d[0]/tot*100:5.1f}% inst {d[1]/tots*100:5.1f}% samp  {d[2]}:{d[3]}: {d[4][:100]}")
