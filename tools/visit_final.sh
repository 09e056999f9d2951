#!/bin/bash
# End-of-round evidence visit (one B200): full -m gpu suite, smoke, bench (both arms, stage times), cfg1 preset,
# GroupNorm / bias+residual microbenchmark, ncu launch list of a 12-frame step.   usage: tools/visit_final.sh [TAG]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-final}
( timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
grep -E "cfg1:|fp16:|passed|failed|rror|exit=" gpurun_out/pytest_gpu_$T.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; tail -1 gpurun_out/smoke_$T.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_$T.csv &
SMI=$!
timeout 900 python bench.py --stage-times > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 3500 gpurun_out/bench_$T.json; tail -2 gpurun_out/bench_$T.err
kill $SMI
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; tail -c 800 gpurun_out/bench_ref_$T.json
timeout 300 python bench.py --preset cfg1 --steps 3 --warmup 2 --no-cpu-baseline --gpu-reference off > gpurun_out/bench_cfg1_$T.json 2> gpurun_out/bench_cfg1_$T.err; tail -c 600 gpurun_out/bench_cfg1_$T.json
timeout 300 python tools/microbench.py gn > gpurun_out/mb_gn_$T.log 2>&1; cut -c1-200 gpurun_out/mb_gn_$T.log
if [ "${RD_SKIP_NCU:-0}" != "1" ]; then
CMD="python bench.py --frames 12 --steps 1 --warmup 1 --no-cpu-baseline --no-output-check --gpu-reference off --coalign-its 200"
timeout 300 $CMD > gpurun_out/plain_$T.log 2>&1 &&
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_launches_$T.log 2>&1
echo "launch list exit=$?"; wc -l gpurun_out/launches_$T.csv
fi
ls gpurun_out | grep "_$T" | tr '\n' ' '
