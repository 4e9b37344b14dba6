"""Parse the ncu --csv metric log written by the candidate / k_grid probes (gpurun_out/cand_probe.csv): one dict of metrics per profiled launch."""
import csv
rows=[r for r in csv.reader(open("gpurun_out/cand_probe.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii],r[ki][:40]),{})[r[mi]]=r[vi]
for k,v in cur.items(): print(k, v)
