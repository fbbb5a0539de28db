#!/usr/bin/env bash
# Multi-GPU session for the headline AND the multi-GPU BASELINE configs (4: upwind5 strong-scaled, 5: fast-waves weak-scaled).
# Charged N x box time: keep it short.  From the repo root:
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_round_multi_w.sh r02q 8 2>&1 | tail -40'
set -u
TAG="${1:-rXX}"; N="${2:-2}"
OUT=gpurun_out; mkdir -p "$OUT"
PORT=29561
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
    print(json.dumps({'value': d['value'], 'ms_per_step': d['ms_per_step'], 'kernel_ms': d['roofline']['kernel_ms'], 'frac': d['roofline']['frac'], 'parallelism': str(c.get('parallelism', c.get('exchange')))[:90],
                      'check': c.get('multi_gpu_check'), 'exposed_us': c.get('exposed_comm_us_per_step')}))
" 2>/dev/null || tail -5 "$OUT/${TAG}_n${N}_${name}.log" | cut -c1-400
}
run 300 bench     bench.py --gpus "$N" --steps 50 --warmup 5
run 200 upwind5   tools/bench_workloads.py --workload upwind5 --gpus "$N" --peer --steps 20
run 200 fastwaves tools/bench_workloads.py --workload fastwaves --gpus "$N" --peer --steps 10
run 200 fastwaves_nccl tools/bench_workloads.py --workload fastwaves --gpus "$N" --steps 10
