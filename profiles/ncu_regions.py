import csv, os, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
agg=collections.Counter(); samp=collections.Counter(); seen=set(); i=0
while i<len(rows):
    r=rows[i]
    if len(r)==2 and r[0]=="File Path":
        f=os.path.basename(r[1]); func=rows[i+1][1]
        if (f,func) in seen: break
        seen.add((f,func)); hdr=rows[i+2]; ci={}
        for k,h in enumerate(hdr): ci.setdefault(h,k)
        i+=3
        while i<len(rows) and not (len(rows[i])==2 and rows[i][0]=="File Path"):
            q=rows[i]
            if len(q)>ci["Instructions Executed"] and q[2]=="-":
                agg[(f,int(q[0]))]+=float(q[ci["Instructions Executed"]] or 0); samp[(f,int(q[0]))]+=float(q[ci["# Samples"]] or 0)
            i+=1
        continue
    i+=1
tot=sum(agg.values()); ts=sum(samp.values())
srcf=sys.argv[2]; src=open(srcf).read().splitlines(); base=os.path.basename(srcf)
marks=[]
for spec in sys.argv[3:]:
    name,pat=spec.split("=",1)
    for k,l in enumerate(src):
        if pat in l: marks.append((name,k+1)); break
marks.append(("end",10**6))
print("total %.2fG"%(tot/1e9))
for (name,a),(_,b) in zip(marks,marks[1:]):
    v=sum(x for (f,l),x in agg.items() if f==base and a<=l<b); s=sum(x for (f,l),x in samp.items() if f==base and a<=l<b)
    print("%-18s inst %5.1f%% (%.2fG) samples %5.1f%%"%(name,100*v/tot,v/1e9,100*s/ts))
v=sum(x for (f,l),x in agg.items() if f!=base); print("other files %.1f%% %.2fG"%(100*v/tot,v/1e9))
