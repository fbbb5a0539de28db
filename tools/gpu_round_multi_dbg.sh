#!/usr/bin/env bash
# forced-schedule runs of bench.py at N ranks (debugging aid for the peer-memory exchange)
set -u
TAG="${1:-rXX}"; N="${2:-2}"
OUT=gpurun_out; mkdir -p "$OUT"
PORT=29641
run() {
  local t="$1" name="$2"; shift 2
  echo "=== $name x$N"; local t0=$SECONDS
  timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$PORT" "$@" > "$OUT/${TAG}_n${N}_${name}.log" 2>&1
  echo "    exit $? after $((SECONDS - t0))s"; PORT=$((PORT + 1))
  grep -h '"metric"' "$OUT/${TAG}_n${N}_${name}.log" | python -c "
import json, sys
for l in sys.stdin:
    d = json.loads(l); c = d['config']
    print(json.dumps({'value': d['value'], 'ms_per_step': d['ms_per_step'], 'kernel_ms': d['roofline']['kernel_ms'], 'opts': c['codegen_options'], 'trial': c.get('schedule_trial'), 'check': c.get('multi_gpu_check')}))
" 2>/dev/null || tail -5 "$OUT/${TAG}_n${N}_${name}.log" | cut -c1-400
}
run 200 peer_tuned   bench.py --gpus "$N" --steps 50 --warmup 5 --step-mode peer
run 200 peer_default bench.py --gpus "$N" --steps 50 --warmup 5 --step-mode peer --no-autotune
GT4PY_B200_SPECIALIZE=off run 200 peer_generic bench.py --gpus "$N" --steps 50 --warmup 5 --step-mode peer --no-autotune
