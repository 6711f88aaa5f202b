set -u
mkdir -p gpurun_out/r02
timeout 300 python tools/ce_probe.py small > gpurun_out/r02/ce_probe2.txt 2>&1; grep -A6 "N=2048\|N=12666\|N=500 " gpurun_out/r02/ce_probe2.txt
cat > /tmp/p.py <<'PY'
import sys; sys.argv=['x']
sys.path.insert(0,'tools')
import ce_probe
ce_probe.run(12666,512,10000)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/ce_launches.csv python /tmp/p.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02/ce_launches.csv')) if len(r)>10 and r[0].isdigit()]
# keep last occurrence set: print the last 40 launches
for r in rows[-46:]:
    print(r[4][:110], r[-1])
PY
