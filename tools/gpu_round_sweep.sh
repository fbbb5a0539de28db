#!/usr/bin/env bash
# Variant sweeps (tools/bench_tma.py) + ncu of the best hdiff variant + one bench.py run.  From the repo root:
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_round_sweep.sh r02k 2>&1 | tail -90'
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
for w in hdiff upwind5 pgrad div; do
  step 300 sweep_$w python tools/bench_tma.py --workload $w
  head -10 "$OUT/${TAG}_sweep_$w.log" | cut -c1-220
done
BEST=$(python - "$OUT/${TAG}_sweep_hdiff.log" <<'PY'
import json, sys
best = None
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    if "options" in d and (best is None or d["ms"] < best["ms"]): best = d
print(json.dumps(best["options"]) if best else '{"interior_loop": true, "static_pitch": 1056}')
PY
)
echo "best hdiff variant: $BEST"
step 300  ncu_best  ncu --set full --clock-control none --import-source on -k regex:b200_hdiff_f32_stream0 -s 30 -c 1 -f -o "$OUT/${TAG}_hdiff_best" \
                    python tools/quick_bench.py --only "$BEST"
step 420  bench     python bench.py --steps 50 --warmup 5
grep -h '"metric"' "$OUT/${TAG}_bench.log" | cut -c1-2500
