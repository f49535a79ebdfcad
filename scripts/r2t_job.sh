set -x
mkdir -p gpurun_out
python scripts/prof_msm.py 21 3 > gpurun_out/r2t_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2t.csv python scripts/prof_msm.py 21 3 > gpurun_out/ncu_r2t.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_r2t.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]; h=rows[hdr]
ki,vi,ui=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Unit")
seq=[]
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    name=r[ki].split("(")[0].replace("void ","").replace("<unnamed>::","").split("<")[0]
    v=float(r[vi].replace(",",""))*{"ns":1e-3,"us":1.0,"ms":1e3}.get(r[ui],1.0)
    seq.append((name,v))
# last MSM = from the last msm_coarse_hist to the end
idx=[i for i,(n,_) in enumerate(seq) if n=="msm_coarse_hist_kernel"][-1]
tot=0
for n,v in seq[idx:]:
    print(f"{n:30s} {v:9.1f} us"); tot+=v
print("sum", tot)
PY
