set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v11
( timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "phase_shuffle or subpixel or upsample" > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
tail -5 gpurun_out/pytest_gpu_$T.log
timeout 600 python tools/microbench.py upconv > gpurun_out/mb_upconv_$T.log 2>&1; cut -c1-700 gpurun_out/mb_upconv_$T.log
for SP in 0 1 0 1; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --gpu-reference off --stage-times --no-e2e --subpixel $SP > gpurun_out/bench_sp${SP}_$T.json 2> gpurun_out/bench_sp${SP}_$T.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_sp${SP}_$T.json').read().strip().splitlines()[-1])
print('subpixel', $SP, round(d['value'],2), d['stage_ms_last_step_rank0'], d['output_check'])
PY
done
