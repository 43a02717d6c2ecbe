#!/bin/bash
# round 2, final evidence on one GPU: launch list of the bench command, ncu --set full of the dominant kernel on both
# bench workloads (tetrahedra 10^6, hexahedra 128^3), traffic JSON, copies under profiles/ are made by the caller
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tet1m.csv \
    python bench.py --steps 2 --warmup 1 --no-krylov > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_tet1m.csv")) if len(r) > 5]
h = rows[0]; tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    d = dict(zip(h, r))
    try: v = float(d["Metric Value"].replace(",", ""))
    except Exception: continue
    u = d.get("Metric Unit", "ns")
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(u, 1e-6)
    k = d["Kernel Name"][:70]; tot[k] += v; cnt[k] += 1
s = sum(tot.values())
for k, v in tot.most_common(12): print(f"{v:10.3f} ms {100*v/s:5.1f} %  x{cnt[k]:3d}  {k}")
PY
# ncu prints template arguments as "(int)7": the regex of the round's earlier captures
timeout 900 bash tools/ncu_capture.sh mpsa_tet1m_r02 'mpsa_kernel.*TileGJ<.int.7,..int.2,..int.24' tet1m 1 | tail -3
timeout 900 bash tools/ncu_capture.sh mpsa_cart128_r02 'mpsa_kernel.*TileGJ<.int.2,..int.3,..int.8' cart128 1 | tail -3
python tools/traffic_json.py gpurun_out/r02_traffic.json tet1m=gpurun_out/mpsa_tet1m_r02_raw.csv cart128=gpurun_out/mpsa_cart128_r02_raw.csv | head -40
rm -f gpurun_out/mpsa_tet1m_r02_source.csv gpurun_out/mpsa_cart128_r02_source.csv
