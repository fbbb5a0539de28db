#!/usr/bin/env bash
# Multi-GPU part of a round's device session (charged N x box time: keep it short).  From the repo root:
#
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_round_multi.sh r02 2 2>&1 | tail -40'
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_round_multi.sh r02 8 2>&1 | tail -40'
set -u
TAG="${1:-rXX}"; N="${2:-2}"
OUT=gpurun_out; mkdir -p "$OUT"
PORT=29541
run() {  # run <seconds> <name> <script and args...>
  local t="$1" name="$2"; shift 2
  echo "=== $name x$N (limit ${t}s)"; local t0=$SECONDS
  timeout "$t" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$PORT" "$@" \
      > "$OUT/${TAG}_n${N}_${name}.log" 2>&1
  echo "    exit $? after $((SECONDS - t0))s"; PORT=$((PORT + 1))
  grep -h '"metric"' "$OUT/${TAG}_n${N}_${name}.log" | cut -c1-700
}
python - <<'PY'
import __graft_entry__ as g
g.build()
PY
run 300 bench        bench.py --gpus "$N" --steps 50 --warmup 5              # schedule trial: serial / overlap / thin
run 200 bench_serial bench.py --gpus "$N" --steps 50 --warmup 5 --step-mode serial --no-autotune
run 200 bench_thin   bench.py --gpus "$N" --steps 50 --warmup 5 --step-mode thin --no-autotune
run 200 exchange     tools/bench_exchange.py
run 200 x2           tools/bench_workloads.py --workload hdiff_x2 --steps 20          # two 2-row exchanges per pass
run 200 x2_fused     tools/bench_workloads.py --workload hdiff_x2 --fuse --steps 20   # one 4-row exchange per pass
run 200 cfg4         tools/bench_workloads.py --workload upwind5 --steps 20
run 200 cfg5         tools/bench_workloads.py --workload fastwaves --steps 10
