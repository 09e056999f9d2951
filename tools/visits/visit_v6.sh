set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v6
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider -k "coalign or bias_residual or groupnorm or fused_modules or output_modes or fixture" > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
grep -E "passed|failed|rror|exit=|^FAILED" gpurun_out/pytest_gpu_$T.log | tail -15
for st in 0 1; do
RD_GN_STAGED=$st timeout 300 python tools/microbench.py gn 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('staged $st', d['bench'], d['case'], round(d['ms'],3), 'ms', round(d['GBs'],0), 'GB/s', round(d['frac_of_hbm'],3))
    except Exception: print(l.strip()[:200])
" | grep groupnorm
done
for pd in 0 1; do
RD_COALIGN_PDL=$pd timeout 200 python tools/microbench.py align --iters 400 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d.get('bench')=='coalign': print('pdl $pd', d['case'], round(d['us_per_iter'],2), 'us/iter', round(d['frac_of_hbm'],3), 'loss', round(d['lossN'],6))
"
done
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --gpu-reference off --stage-times > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$T.json').read().strip().splitlines()[-1])
print('bench fast', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'attn', round(d['roofline']['achieved'],1), d['stage_ms_last_step_rank0'], d['output_check'])
PY
timeout 600 python bench.py --preset full --steps 1 --warmup 1 --no-cpu-baseline --gpu-reference off --stage-times > gpurun_out/bench_full_1gpu_$T.json 2> gpurun_out/bench_full_1gpu_$T.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full_1gpu_$T.json').read().strip().splitlines()[-1])
print('bench full', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'attn', round(d['roofline']['achieved'],1), d['roofline']['share_of_step'], d['stage_ms_last_step_rank0'], d['output_check'])
PY
