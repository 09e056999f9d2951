#!/bin/bash
# Multi-GPU visit: tools/visit_multi.sh TAG NGPU [tests] [fast] [fast_ar] [full] [full1]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=$1; N=$2; shift 2
ALL="$*"
has() { case " $ALL " in *" $1 "*) return 0;; esac; return 1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get('roofline') or {}
    print(sys.argv[1], 'N', d['n_gpus'], 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/step', round(d['ms_per_step'],1),
          'attn', round(r.get('achieved',0),1), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'check', (d.get('output_check') or {}).get('ok'))
    print('   stages', d.get('stage_ms_last_step_rank0'))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
}
if has mtest; then
  ( timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -k "[$N]" > gpurun_out/pytest_multi_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_multi_$T.log )
  grep -E "passed|failed|rror|exit=" gpurun_out/pytest_multi_$T.log | tail -5
fi
if has tests; then
  ( timeout 1500 python -m pytest tests/test_gpu_pipeline.py tests/test_video_call.py tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_multi_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_multi_$T.log )
  grep -E "cfg1:|fp16:|passed|failed|rror|exit=" gpurun_out/pytest_multi_$T.log | tail -12
fi
if has fast; then
  timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 2 --stage-times > gpurun_out/bench_fast_${N}gpu_$T.json 2> gpurun_out/bench_fast_${N}gpu_$T.err; summ gpurun_out/bench_fast_${N}gpu_$T.json; tail -2 gpurun_out/bench_fast_${N}gpu_$T.err
fi
if has fast_ar; then
  timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 2 --stage-times --coalign-mode allreduce > gpurun_out/bench_fast_ar_${N}gpu_$T.json 2> gpurun_out/bench_fast_ar_${N}gpu_$T.err; summ gpurun_out/bench_fast_ar_${N}gpu_$T.json; tail -2 gpurun_out/bench_fast_ar_${N}gpu_$T.err
fi
if has full; then
  timeout 900 $TR bench.py --gpus $N --preset full --steps 1 --warmup 1 --stage-times > gpurun_out/bench_full_${N}gpu_$T.json 2> gpurun_out/bench_full_${N}gpu_$T.err; summ gpurun_out/bench_full_${N}gpu_$T.json; tail -2 gpurun_out/bench_full_${N}gpu_$T.err
fi
if has fast1; then
  timeout 600 python bench.py --steps 3 --warmup 2 --stage-times --no-cpu-baseline --gpu-reference off > gpurun_out/bench_fast_1gpu_$T.json 2> gpurun_out/bench_fast_1gpu_$T.err; summ gpurun_out/bench_fast_1gpu_$T.json
fi
if has full1; then
  timeout 900 python bench.py --preset full --steps 1 --warmup 1 --stage-times --no-cpu-baseline --gpu-reference off > gpurun_out/bench_full_1gpu_$T.json 2> gpurun_out/bench_full_1gpu_$T.err; summ gpurun_out/bench_full_1gpu_$T.json; tail -2 gpurun_out/bench_full_1gpu_$T.err
fi
