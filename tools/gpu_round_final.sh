#!/usr/bin/env bash
# Last device session of a round: smoke + the whole GPU suite on the final tree, then the end-to-end call at several K-slab counts.
#   /usr/local/graft/bin/gpurun --timeout 700 -- 'bash tools/gpu_round_final.sh r02t 2>&1 | tail -40'
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
step 120 smoke python __graft_entry__.py smoke
step 400 tests_gpu python -m pytest tests -q -m gpu -x --durations=3
tail -2 "$OUT/${TAG}_tests_gpu.log"
WIN='{"interior_loop": true, "static_pitch": 1056, "tma": 3, "tile_j": 32, "prefetch": 1, "tma_mode": "bulk", "stcs": true}'
for C in 10 20 32 40; do
  GT4PY_B200_HOST_CHUNKS=$C step 120 e2e_$C python bench.py --steps 20 --warmup 5 --no-cpu-baseline --options "$WIN" --pipeline-chunks $C
  python - "$OUT/${TAG}_e2e_$C.log" $C <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        d = json.loads(l); e = d["e2e"]
        print("chunks", sys.argv[2], "value", d["value"], "e2e", e["value"], "hostpipe", e.get("hostpipe_value"), "serial", e.get("serial_value"), e.get("note"))
PY
done
