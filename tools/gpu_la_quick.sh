# linear attention: tests + stand-alone timing of both output-pass versions + launch list of v2
mkdir -p gpurun_out; P=gpurun_out/${1:-laq}
timeout 240 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "linattn or forward_layerwise" > ${P}_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 ${P}_pytest.log
timeout 120 python tools/run_linattn.py 2>&1 | tee ${P}_plain.log
IDIFF_LA_OUT_V1=1 timeout 120 python tools/run_linattn.py 2>&1 | tee ${P}_plain_v1.log
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'la_out' --csv --log-file ${P}_launches.csv python tools/run_linattn.py > /dev/null 2>&1; echo "list exit=$?"
python - <<PY
import csv,re
lines=[l for l in open('${P}_launches.csv') if l.startswith('"')]
r=csv.reader(lines); hdr=next(r)
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size'); mi=hdr.index('Metric Name')
acc={}
for row in r:
    m=re.search(r'(la_\w+<\d+>)',row[ki])
    if m: acc.setdefault((m.group(1),row[gi],row[mi]),[]).append(float(row[vi].replace(',','')))
for k,v in acc.items(): print(k, len(v), sorted(v)[len(v)//2])
PY
