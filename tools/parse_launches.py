import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
h=rows[hdr]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
out=[]
for r in rows[hdr+1:]:
    if len(r)>vi: out.append((r[ki].replace("<unnamed>::","").replace("void ","")[:34], float(r[vi])/1000))
# last re-bin = kernels after the last step_kernel
idx=max(i for i,(k,_) in enumerate(out) if k.startswith("step_kernel"))
reb=out[idx+1:]
print("re-bin:", " ".join(f"{k.split('(')[0]}={v:.0f}" for k,v in reb), "| total %.0f us" % sum(v for _,v in reb))
first=[i for i,(k,_) in enumerate(out) if k.startswith("key_count")][1]
ing=out[first:idx-3] if False else out[first:[i for i,(k,_) in enumerate(out) if k.startswith("step_kernel")][0]]
print("ingest:", " ".join(f"{k.split('(')[0]}={v:.0f}" for k,v in ing), "| total %.0f us" % sum(v for _,v in ing))
