#!/usr/bin/env bash
# Device session for the bulk-async (TMA) variant of the streaming kernel.  From the repo root on the GPU box:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round_tma.sh r02c 2>&1 | tail -80'
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
  step 300 tma_$w python tools/bench_tma.py --workload $w
  head -8 "$OUT/${TAG}_tma_$w.log" | cut -c1-200
done
BEST=$(python - "$OUT/${TAG}_tma_hdiff.log" <<'PY'
import json, sys
best = None
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    if "options" in d and d["options"].get("tma") and (best is None or d["ms"] < best["ms"]): best = d
print(json.dumps(best["options"]) if best else '{"interior_loop": true, "static_pitch": 1056, "tma": 2}')
PY
)
echo "best bulk-async variant: $BEST"
step 300  ncu_tma  ncu --set full --clock-control none --import-source on -k regex:b200_hdiff_f32_stream0 -s 30 -c 1 -f -o "$OUT/${TAG}_hdiff_tma" \
                   python tools/quick_bench.py --only "$BEST"
