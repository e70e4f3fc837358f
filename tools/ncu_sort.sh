M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_atom.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum
for v in $VLIST; do
VARIANTS=$v ncu --metrics $M --clock-control none -k regex:sort -c 6 --csv --log-file gpurun_out/sort_v$v.csv python tools/exp_sort.py > /dev/null 2>&1
python - <<PY
import csv,io
txt=open("gpurun_out/sort_v$v.csv").read()
rows=list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
cur={}
for r in rows:
    k=(r["ID"],r["Kernel Name"][:44],r["Grid Size"])
    cur.setdefault(k,{})[r["Metric Name"].split("__")[-1][:28]]=r["Metric Value"]
for k,v in cur.items(): print(k, v)
PY
done
