#!/bin/bash
# One GPU-box visit for the round record: full -m gpu suite, smoke, bench.py (both arms), then the ncu passes
# (launch list of a short-clip bench step; --set full of the attention and co-alignment kernels).
# Everything lands in gpurun_out/; tools/ncu_summary.py + tools/launch_shares.py turn it into profiles/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-r1}
( timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log ) 
tail -3 gpurun_out/pytest_gpu_$T.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; tail -1 gpurun_out/smoke_$T.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_$T.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 3000 gpurun_out/bench_$T.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; tail -c 1500 gpurun_out/bench_ref_$T.json
kill $SMI
timeout 600 python tools/microbench.py attn > gpurun_out/mb_attn_$T.log 2>&1; tail -8 gpurun_out/mb_attn_$T.log
timeout 600 python tools/microbench.py align --iters 200 > gpurun_out/mb_align_$T.log 2>&1; tail -6 gpurun_out/mb_align_$T.log
timeout 300 python tools/microbench.py gn > gpurun_out/mb_gn_$T.log 2>&1; tail -6 gpurun_out/mb_gn_$T.log
if [ "${RD_SKIP_NCU:-0}" != "1" ]; then
CMD="python bench.py --frames 12 --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_$T.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_launches_$T.log 2>&1
echo "launch list exit=$?"; wc -l gpurun_out/launches_$T.csv
CMD2=(python tools/microbench.py attn --only "full L0")
timeout 300 "${CMD2[@]}" > gpurun_out/plain2_$T.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:xattn_fwd -s 3 -c 1 -f -o gpurun_out/prof_xattn_$T "${CMD2[@]}" > gpurun_out/ncu_xattn_$T.log 2>&1
echo "ncu xattn exit=$?"
CMD3="python tools/microbench.py align --only 1024x576 --iters 20"
timeout 300 $CMD3 > gpurun_out/plain3_$T.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coalign_pass -s 8 -c 1 -f -o gpurun_out/prof_coalign_$T $CMD3 > gpurun_out/ncu_coalign_$T.log 2>&1
echo "ncu coalign exit=$?"
fi
ls -la gpurun_out | tail -20
