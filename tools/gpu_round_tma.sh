#!/usr/bin/env bash
# Device session for the bulk-async (TMA) variant of the streaming kernel.  From the repo root on the GPU box:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round_tma.sh r02f 2>&1 | tail -90'
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
step 1500 tests_gpu python -m pytest tests -q -m gpu --durations=5
tail -12 "$OUT/${TAG}_tests_gpu.log" | cut -c1-250
for w in hdiff upwind5 pgrad div; do
  step 300 tma_$w python tools/bench_tma.py --workload $w
  head -8 "$OUT/${TAG}_tma_$w.log" | cut -c1-200
done
BEST=$(python - "$OUT/${TAG}_tma_hdiff.log" <<'PY'
import json, sys
best = None
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    if "options" in d and d["options"].get("tma") and d["options"].get("tma_mode") != "bulk" and (best is None or d["ms"] < best["ms"]): best = d
print(json.dumps(best["options"]) if best else '{"interior_loop": true, "static_pitch": 1056, "tma": 3}')
PY
)
echo "best tensor-map variant: $BEST"
step 300  ncu_tma  ncu --set full --clock-control none --import-source on -k regex:b200_hdiff_f32_stream0 -s 30 -c 1 -f -o "$OUT/${TAG}_hdiff_tma" \
                   python tools/quick_bench.py --only "$BEST"
step 300  bench_ref  python bench.py --impl reference --steps 20 --warmup 5
step 420  bench      python bench.py --steps 50 --warmup 5
grep -h '"metric"' "$OUT/${TAG}_bench_ref.log" "$OUT/${TAG}_bench.log" | cut -c1-1800
