#!/usr/bin/env bash
# Session r02p: defaults after the r02n/r02o sweeps (column look-ahead rule, prefetch rule), every BASELINE config with
# the per-stencil autotuner, bench.py with launch list + full ncu capture of the variant it timed.  From the repo root:
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_round_p.sh r02p 2>&1 | tail -120'
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
step 600 tests_gpu python -m pytest tests -q -m gpu -x --durations=3
tail -3 "$OUT/${TAG}_tests_gpu.log"
CB='[{}, {"seq_rotate": false}, {"seq_rotate": true}, {"seq_rotate": true, "seq_prefetch": 2}, {"seq_prefetch": 2}, {"seq_prefetch": 4}, {"fuse_columns": true}]'
for c in "tridiagonal_f64 512,512,160" "fw_wsolve_f32 4096,512,80" "vadv_f64 512,512,160"; do
  set -- $c
  step 200 cols_$1 python tools/quick_bench.py --name $1 --variant default --domain $2 --candidates "$CB"
  cut -c100-330 "$OUT/${TAG}_cols_$1.log"
done
step 300 cfg3 python tools/bench_workloads.py --workload tridiagonal --steps 20
step 300 cfg4 python tools/bench_workloads.py --workload upwind5 --steps 20 --tune
step 300 cfg4_default python tools/bench_workloads.py --workload upwind5 --steps 20
step 400 cfg5 python tools/bench_workloads.py --workload fastwaves --steps 10 --tune
step 300 cfg5_default python tools/bench_workloads.py --workload fastwaves --steps 10
step 300 x2 python tools/bench_workloads.py --workload hdiff_x2 --steps 20
step 300 x2_fused python tools/bench_workloads.py --workload hdiff_x2 --fuse --steps 20
step 420 bench python bench.py --steps 50 --warmup 5
WIN=$(python - "$OUT/${TAG}_bench.log" <<'PY'
import json, sys
opts = {"interior_loop": True, "static_pitch": 1056}
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        opts = {k: v for k, v in json.loads(l)["config"]["codegen_options"].items() if k not in ("strategy", "device_sync")}
print(json.dumps(opts))
PY
)
echo "bench selected: $WIN"
echo "$WIN" > "$OUT/${TAG}_bench_variant.json"
step 300 ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${TAG}_launches_bench.csv" \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --options "$WIN"
step 300 ncu_hdiff ncu --set full --clock-control none --import-source on -k regex:b200_hdiff_f32_stream0 -s 8 -c 1 -f -o "$OUT/${TAG}_hdiff_selected" \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline --options "$WIN"
grep -h '"metric"' "$OUT/${TAG}"_bench.log "$OUT/${TAG}"_cfg*.log "$OUT/${TAG}"_x2*.log 2>/dev/null | cut -c1-1200
