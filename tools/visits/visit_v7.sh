set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v7
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider -k "softmax or layernorm or vae_mid or fused_modules or output_modes or fixture or fp16 or coalign_free" > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
grep -E "passed|failed|rror|exit=|^FAILED" gpurun_out/pytest_gpu_$T.log | tail -15
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --gpu-reference off --stage-times > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$T.json').read().strip().splitlines()[-1])
print('bench fast', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'attn', round(d['roofline']['achieved'],1), d['stage_ms_last_step_rank0'], d['output_check'])
PY
tail -3 gpurun_out/bench_$T.err
