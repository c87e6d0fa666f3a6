import csv,sys,subprocess,io,collections
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr=None; data=[]
for r in rows:
    if 'Source' in r and '# Samples' in ' '.join(r): hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(r)
ix={h:i for i,h in enumerate(hdr)}
S=ix['# Samples']; E=ix['Instructions Executed']
tot=sum(float(r[S] or 0) for r in data)
print("total samples",tot,"rows",len(data))
print("columns:", [h for h in hdr if 'stall' in h.lower() or 'Stall' in h][:6])
top=sorted(range(len(data)), key=lambda i:-float(data[i][S] or 0))[:60]
for i in sorted(top):
    r=data[i]
    print("%5d %-60s smp %5.2f%% exec %12.0f | %s"%(i, r[ix['Source']].strip()[:60], 100*float(r[S])/tot, float(r[E] or 0), r[ix['Warp Stall Sampling (All Samples)']][:100] if 'Warp Stall Sampling (All Samples)' in ix else ''))
# opcode totals
c=collections.Counter()
for r in data:
    src=r[ix['Source']].strip()
    if not src: continue
    op=src.split()[1] if src.startswith('@') else src.split()[0]
    c[op.split('.')[0]]+=float(r[S] or 0)
print({k:round(100*v/tot,1) for k,v in c.most_common(14)})
