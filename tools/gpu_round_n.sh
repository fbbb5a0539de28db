#!/usr/bin/env bash
# Session r02n: the round-2 generator changes on the device (divisions by launch invariants, I-cache-aware unroll factor,
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
for w in upwind5 pgrad div hdiff; do
  step 240 sweep_$w python tools/bench_tma.py --workload $w --candidates tools/cands_r02n.json
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
COLS='[{}, {"fuse_columns": true}, {"col_smem": true}, {"col_smem": true, "seq_prefetch": 2}, {"col_smem": true, "seq_prefetch": 4}, {"col_smem": true, "seq_prefetch": 6}, {"col_smem": true, "seq_prefetch": 8}, {"col_smem": true, "seq_prefetch": 4, "col_smem_block": [64, 2]}, {"col_smem": true, "seq_prefetch": 8, "col_smem_block": [64, 2]}, {"col_smem": true, "seq_prefetch": 4, "col_smem_block": [32, 1]}, {"col_smem": true, "seq_prefetch": 8, "col_smem_block": [32, 1]}, {"col_smem": true, "seq_prefetch": 12, "col_smem_block": [32, 1]}, {"div_inv": false}]'
step 300 cols_wsolve python tools/quick_bench.py --name fw_wsolve_f32 --variant default --domain 4096,512,80 --candidates "$COLS"
cut -c1-400 "$OUT/${TAG}_cols_wsolve.log"
step 200 cols_vadv python tools/quick_bench.py --name vadv_f64 --variant default --domain 512,512,40 --candidates '[{}, {"fuse_columns": true}, {"col_smem": true, "col_smem_kb": 100}, {"col_smem": true, "col_smem_kb": 100, "seq_prefetch": 4}]'
cut -c1-400 "$OUT/${TAG}_cols_vadv.log"
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
step 240 ncu_wsolve ncu --set full --clock-control none --import-source on -k regex:wsolve -s 12 -c 1 -f -o "$OUT/${TAG}_wsolve_smem" \
     python tools/quick_bench.py --name fw_wsolve_f32 --variant default --domain 4096,512,80 --iters 3 --only '{"col_smem": true, "seq_prefetch": 4}'
step 300 cfg4 python tools/bench_workloads.py --workload upwind5 --steps 20
step 300 cfg5 python tools/bench_workloads.py --workload fastwaves --steps 10
grep -h '"metric"' "$OUT/${TAG}"_cfg*.log 2>/dev/null | cut -c1-900
