#!/usr/bin/env bash
# One gpurun call that measures everything a round needs, most valuable first, every step under its own
# timeout so that a hang cannot eat the box (run from the repo root on the GPU box):
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02 2>&1 | tail -70'
#
# Outputs land in gpurun_out/<tag>_* (merged back by gpurun); summaries for profiles/ are produced HERE
# afterwards with tools/ncu_summary.py.  Numbers printed under ncu are never bench values.
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
step 120  smoke      python __graft_entry__.py smoke
step 300  bench_ref  python bench.py --impl reference --steps 20 --warmup 5
step 420  bench      python bench.py --steps 50 --warmup 5
# the variant the bench selected, as a JSON dict for --options
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
step 900  tests_gpu  python -m pytest tests -q -m gpu --durations=5
tail -4 "$OUT/${TAG}_tests_gpu.log"
# launch list of the bench command with the selected variant (per-launch times are cold-cache / serialised: shares, not absolutes)
step 300  ncu_list   ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${TAG}_launches_bench.csv" \
                     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --options "$WIN"
# full capture of the selected variant inside bench.py's stepping loop (launch 8 of that name: past the first-call work)
step 300  ncu_hdiff  ncu --set full --clock-control none --import-source on -k regex:b200_hdiff_f32_stream0 -s 8 -c 1 -f -o "$OUT/${TAG}_hdiff_selected" \
                     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pipeline --options "$WIN"
# the other BASELINE configs (single GPU) with the variants the sweeps found (profiles/README.md)
step 300  cfg3       python tools/bench_workloads.py --workload tridiagonal --steps 20
step 300  cfg4       python tools/bench_workloads.py --workload upwind5 --steps 20 --tune
step 400  cfg5       python tools/bench_workloads.py --workload fastwaves --steps 10 --tune
step 300  x2         python tools/bench_workloads.py --workload hdiff_x2 --steps 20
step 300  x2_fused   python tools/bench_workloads.py --workload hdiff_x2 --fuse --steps 20
step 300  ncu_tri    ncu --set full --clock-control none --import-source on -k regex:tridiagonal -c 2 -f -o "$OUT/${TAG}_tridiagonal" \
                     python tools/bench_workloads.py --workload tridiagonal --steps 2 --warmup 3
step 300  ncu_up5    ncu --set full --clock-control none --import-source on -k regex:upwind5 -s 4 -c 1 -f -o "$OUT/${TAG}_upwind5" \
                     python tools/bench_workloads.py --workload upwind5 --steps 2 --warmup 3
step 300  ncu_fw     ncu --set full --clock-control none --import-source on -k regex:b200_fw_ -s 12 -c 6 -f -o "$OUT/${TAG}_fastwaves" \
                     python tools/bench_workloads.py --workload fastwaves --steps 2 --warmup 3
step 300  ncu_fused  ncu --set full --clock-control none --import-source on -k regex:fused2_stream0 -s 4 -c 1 -f -o "$OUT/${TAG}_hdiff_fused2" \
                     python tools/bench_workloads.py --workload hdiff_x2 --fuse --steps 3 --warmup 3
grep -h '"metric"' "$OUT/${TAG}"_bench*.log "$OUT/${TAG}"_cfg*.log "$OUT/${TAG}"_x2*.log 2>/dev/null | cut -c1-700
