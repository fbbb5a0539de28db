#!/usr/bin/env bash
# Multi-GPU part of a round's device session (charged N x box time: keep it short).  From the repo root:
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_round_multi.sh r02 2 2>&1 | tail -40'
set -u
TAG="${1:-rXX}"; N="${2:-2}"; WHAT="${3:-all}"
OUT=gpurun_out; mkdir -p "$OUT"
PORT=29541
run() {  # run <seconds> <name> <script and args...>
  local t="$1" name="$2"; shift 2
  echo "=== $name x$N (limit ${t}s)"; local t0=$SECONDS
  timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$PORT" "$@" \
      > "$OUT/${TAG}_n${N}_${name}.log" 2>&1
  echo "    exit $? after $((SECONDS - t0))s"; PORT=$((PORT + 1))
  grep -h '"metric"' "$OUT/${TAG}_n${N}_${name}.log" | python -c "
import json, sys
for l in sys.stdin:
    d = json.loads(l); c = d['config']
    print(json.dumps({'value': d['value'], 'ms_per_step': d['ms_per_step'], 'kernel_ms': d['roofline']['kernel_ms'], 'parallelism': c['parallelism'][:90],
                      'trial': c.get('schedule_trial'), 'check': c.get('multi_gpu_check'), 'exposed_us': c.get('exposed_comm_us_per_step'), 'note': d['e2e'].get('note')}))
" 2>/dev/null || tail -5 "$OUT/${TAG}_n${N}_${name}.log" | cut -c1-400
}
run 300 bench        bench.py --gpus "$N" --steps 50 --warmup 5              # schedule trial: serial / overlap / thin / peer
if [ "$WHAT" = "all" ]; then
run 200 bench_peer   bench.py --gpus "$N" --steps 50 --warmup 5 --step-mode peer --no-autotune
run 200 bench_serial bench.py --gpus "$N" --steps 50 --warmup 5 --step-mode serial --no-autotune
fi
