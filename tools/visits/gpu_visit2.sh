#!/bin/bash
# Round-2 GPU-box visit: -m gpu suite, smoke, bench.py (fast, stage times), microbenchmarks, ncu launch list.
# usage: tools/visits/gpu_visit2.sh TAG [tests] [bench] [mb] [launches] [ncu_xattn] [ncu_coalign] [full]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=$1; shift
has() { for a in "$@"; do :; done; case " $ALL " in *" $1 "*) return 0;; esac; return 1; }
ALL="$*"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$T.txt 2>&1
if has tests; then
  ( timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider -s > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
  grep -E "cfg1:|fp16:|passed|failed|error|exit=" gpurun_out/pytest_gpu_$T.log | tail -12
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; tail -1 gpurun_out/smoke_$T.log
fi
if has bench; then
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_$T.csv &
  SMI=$!
  timeout 900 python bench.py --stage-times > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 4500 gpurun_out/bench_$T.json; tail -3 gpurun_out/bench_$T.err
  kill $SMI
fi
if has full; then
  timeout 900 python bench.py --preset full --steps 1 --warmup 1 --stage-times --no-cpu-baseline > gpurun_out/bench_full_$T.json 2> gpurun_out/bench_full_$T.err; tail -c 3000 gpurun_out/bench_full_$T.json; tail -3 gpurun_out/bench_full_$T.err
fi
if has mb; then
  timeout 600 python tools/microbench.py attn > gpurun_out/mb_attn_$T.log 2>&1; tail -14 gpurun_out/mb_attn_$T.log | cut -c1-600
  timeout 600 python tools/microbench.py align --iters 200 > gpurun_out/mb_align_$T.log 2>&1; tail -6 gpurun_out/mb_align_$T.log
  timeout 300 python tools/microbench.py gn > gpurun_out/mb_gn_$T.log 2>&1; tail -6 gpurun_out/mb_gn_$T.log
fi
if has sweep; then
  timeout 900 python tools/microbench.py attn --full > gpurun_out/mb_attn_sweep_$T.log 2>&1; tail -4 gpurun_out/mb_attn_sweep_$T.log | cut -c1-400
fi
if has launches; then
  CMD="python bench.py --frames 30 --steps 1 --warmup 1 --no-cpu-baseline --no-output-check --gpu-reference off"
  timeout 600 $CMD > gpurun_out/plain_$T.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_launches_$T.log 2>&1
  echo "launch list exit=$?"; wc -l gpurun_out/launches_$T.csv
fi
if has ncu_xattn; then
  CMD2=(python tools/microbench.py attn --only "full L0 B=4" --no-sdpa)
  timeout 300 "${CMD2[@]}" > gpurun_out/plain2_$T.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:xattn_fwd -s 3 -c 1 -f -o gpurun_out/prof_xattn_$T "${CMD2[@]}" > gpurun_out/ncu_xattn_$T.log 2>&1
  echo "ncu xattn exit=$?"
fi
if has ncu_coalign; then
  CMD3="python tools/microbench.py align --only 1024x576 --iters 20"
  timeout 300 $CMD3 > gpurun_out/plain3_$T.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:coalign_pass -s 8 -c 1 -f -o gpurun_out/prof_coalign_$T $CMD3 > gpurun_out/ncu_coalign_$T.log 2>&1
  echo "ncu coalign exit=$?"
fi
ls gpurun_out | grep "_$T" | tr '\n' ' '
