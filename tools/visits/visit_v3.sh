set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v3
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_video_call.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider -k "coalign or video or output_modes or merge or min_and" > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
grep -E "passed|failed|rror|exit=|^FAILED" gpurun_out/pytest_gpu_$T.log | tail -15
for c in 0 1 2 3 4; do
RD_COALIGN_CFG=$c timeout 200 python tools/microbench.py align --iters 300 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d.get('bench')=='coalign': print('cfg $c', d['case'], round(d['us_per_iter'],2), 'us/iter', round(d['GBs'],0), 'GB/s', round(d['frac_of_hbm'],3), 'loss', round(d['lossN'],6))
"
done
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --gpu-reference off --stage-times > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$T.json').read().strip().splitlines()[-1])
print('bench', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'attn', round(d['roofline']['achieved'],1), d['stage_ms_last_step_rank0'], d['output_check'])
PY
