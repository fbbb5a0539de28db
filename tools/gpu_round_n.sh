#!/usr/bin/env bash
# Sessions r02n / r02o: the round-2 generator changes on the device (divisions by launch invariants, I-cache-aware unroll factor,
# level-fastest task order, shared-memory temporaries of fused sweeps).  From the repo root:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round_n.sh r02n 2>&1 | tail -120'
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
step() {  # step <seconds> <name> <command...>
  local t="$1" name="$2"; shift 2
  echo "=== $name (limit ${t}s)"; local t0=$SECONDS
  timeout "$t" "$@" > "$OUT/${TAG}_${name}.log" 2>&1
  echo "    exit $? after $((SECONDS - t0))s -> $OUT/${TAG}_${name}.log"
}
step 600 tests_gpu python -m pytest tests -q -m gpu -x --durations=3
tail -3 "$OUT/${TAG}_tests_gpu.log"
for w in $(python -c "import json,sys; print(' '.join(json.load(open('tools/cands_${TAG}.json'))))"); do
  step 240 sweep_$w python tools/bench_tma.py --workload $w --candidates tools/cands_${TAG}.json
  sort -t: -k4 "$OUT/${TAG}_sweep_$w.log" | python -c "
import sys, json
rows = []
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    if 'ms' in d: rows.append(d)
    elif d.get('rejected'): print('REJECTED', d)
for d in sorted(rows, key=lambda d: d['ms']): print(d['ms'], d['frac_of_peak'], json.dumps(d['options']))
"
done
WS='[{}, {"seq_rotate": false}, {"seq_prefetch": 2}, {"seq_prefetch": 4}, {"fuse_columns": true}, {"fuse_columns": true, "seq_prefetch": 2}, {"fuse_columns": true, "seq_prefetch": 4}, {"col_smem": true}, {"col_smem": true, "seq_prefetch": 2}, {"col_smem": true, "seq_prefetch": 4}, {"col_smem": true, "seq_prefetch": 6}, {"col_smem": true, "seq_prefetch": 8}, {"col_smem": true, "seq_prefetch": 4, "col_smem_block": [64, 2]}, {"col_smem": true, "seq_prefetch": 8, "col_smem_block": [64, 2]}, {"col_smem": true, "seq_prefetch": 6, "col_smem_block": [32, 1]}, {"col_smem": true, "seq_prefetch": 6, "col_hints": true}]'
step 300 cols_wsolve python tools/quick_bench.py --name fw_wsolve_f32 --variant default --domain 4096,512,80 --candidates "$WS"
cut -c113-400 "$OUT/${TAG}_cols_wsolve.log"
TR='[{}, {"seq_rotate": false}, {"seq_prefetch": 2}, {"seq_prefetch": 3}, {"seq_prefetch": 4}, {"col_hints": true}, {"col_hints": true, "seq_prefetch": 2}, {"fuse_columns": true}, {"fuse_columns": true, "seq_prefetch": 2}, {"fuse_columns": true, "seq_prefetch": 4}, {"fuse_columns": true, "seq_prefetch": 4, "col_hints": true}, {"fuse_columns": true, "seq_prefetch": 4, "seq_smem_pad": 112640}, {"fuse_columns": true, "seq_prefetch": 4, "seq_smem_pad": 76800}, {"fuse_columns": true, "seq_prefetch": 4, "seq_smem_pad": 57344}, {"fuse_columns": true, "seq_prefetch": 4, "seq_smem_pad": 40960}, {"fuse_columns": true, "seq_prefetch": 6, "seq_smem_pad": 76800, "col_hints": true}, {"fuse_columns": true, "seq_prefetch": 6, "seq_smem_pad": 57344, "col_hints": true}, {"fuse_columns": true, "seq_prefetch": 6, "seq_smem_pad": 40960, "col_hints": true}, {"fuse_columns": true, "seq_prefetch": 8, "seq_smem_pad": 57344, "col_hints": true}, {"fuse_columns": true, "seq_prefetch": 3, "seq_smem_pad": 32768, "col_hints": true}]'
step 300 cols_tri python tools/quick_bench.py --name tridiagonal_f64 --variant default --domain 512,512,160 --candidates "$TR"
cut -c113-400 "$OUT/${TAG}_cols_tri.log"
step 200 cols_vadv python tools/quick_bench.py --name vadv_f64 --variant default --domain 512,512,160 --candidates '[{}, {"seq_rotate": false}, {"seq_prefetch": 2}, {"seq_prefetch": 3}, {"fuse_columns": true}, {"fuse_columns": true, "seq_prefetch": 2}, {"fuse_columns": true, "seq_prefetch": 3}, {"fuse_columns": true, "seq_prefetch": 2, "col_hints": true}]'
cut -c113-400 "$OUT/${TAG}_cols_vadv.log"
bestcol() { python - "$1" <<'PY'
import json, sys
best = None
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    if "ms" in d and (best is None or d["ms"] < best["ms"]): best = d
keys = ("seq_rotate", "seq_prefetch", "fuse_columns", "col_smem", "col_smem_block", "col_hints", "seq_smem_pad")
print(json.dumps({k: best[k] for k in keys if k in best}) if best else "{}")
PY
}
# ncu of the best upwind5 / pgrad variants and of the shared-memory w solver
best() { python - "$1" <<'PY'
import json, sys
best = None
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    if "options" in d and "ms" in d and (best is None or d["ms"] < best["ms"]): best = d
print(json.dumps(best["options"]) if best else "{}")
PY
}
B5=$(best "$OUT/${TAG}_sweep_upwind5.log"); echo "best upwind5: $B5"
step 240 ncu_up5 ncu --set full --clock-control none --import-source on -k regex:upwind5 -s 12 -c 1 -f -o "$OUT/${TAG}_upwind5_best" \
     python tools/quick_bench.py --name upwind5_f32 --domain 2048,2048,80 --iters 3 --only "$B5"
BP=$(best "$OUT/${TAG}_sweep_pgrad.log"); echo "best pgrad: $BP"
step 240 ncu_pgrad ncu --set full --clock-control none --import-source on -k regex:pgrad -s 36 -c 3 -f -o "$OUT/${TAG}_pgrad_best" \
     python tools/quick_bench.py --name fw_pgrad_f32 --domain 4096,512,80 --iters 3 --only "$BP"
BW=$(bestcol "$OUT/${TAG}_cols_wsolve.log"); echo "best wsolve: $BW"
step 240 ncu_wsolve ncu --set full --clock-control none --import-source on -k regex:wsolve -s 12 -c 2 -f -o "$OUT/${TAG}_wsolve_best" \
     python tools/quick_bench.py --name fw_wsolve_f32 --variant default --domain 4096,512,80 --iters 3 --only "$BW"
BT=$(bestcol "$OUT/${TAG}_cols_tri.log"); echo "best tridiagonal: $BT"
step 240 ncu_tri ncu --set full --clock-control none --import-source on -k regex:tridiagonal -s 12 -c 2 -f -o "$OUT/${TAG}_tridiagonal_best" \
     python tools/quick_bench.py --name tridiagonal_f64 --variant default --domain 512,512,160 --iters 3 --only "$BT"
step 300 cfg3 python tools/bench_workloads.py --workload tridiagonal --steps 20
step 300 cfg4 python tools/bench_workloads.py --workload upwind5 --steps 20
step 300 cfg5 python tools/bench_workloads.py --workload fastwaves --steps 10
grep -h '"metric"' "$OUT/${TAG}"_cfg*.log 2>/dev/null | cut -c1-900
