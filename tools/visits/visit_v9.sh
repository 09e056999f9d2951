set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v9
( timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider -k "unload or output_modes" > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
grep -E "passed|failed|rror|exit=|^FAILED" gpurun_out/pytest_gpu_$T.log | tail -6
timeout 300 python bench.py --preset cfg1 --steps 3 --warmup 2 --no-cpu-baseline --gpu-reference off > gpurun_out/bench_cfg1_$T.json 2> gpurun_out/bench_cfg1_$T.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg1_$T.json').read().strip().splitlines()[-1])
print('cfg1', round(d['value'],2), d['output_check'])
PY
timeout 900 python bench.py --preset paper --frames 60 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --gpu-reference off --stage-times --cudnn-benchmark 0 > gpurun_out/bench_paper60_$T.json 2> gpurun_out/bench_paper60_$T.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_paper60_$T.json').read().strip().splitlines()[-1])
print('paper60', round(d['value'],3), 'ms/step', round(d['ms_per_step'],0), 'attn', round(d['roofline']['achieved'],1), d['roofline']['share_of_step'], d['stage_ms_last_step_rank0'], d['output_check'])
PY
tail -3 gpurun_out/bench_paper60_$T.err
