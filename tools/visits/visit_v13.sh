set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v13
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "groupnorm or fused_modules or upsample or phase" > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
tail -4 gpurun_out/pytest_gpu_$T.log
for cfg in "0 3" "1 2" "1 3" "1 4" "1 6"; do
set -- $cfg
echo "== RD_GN_STAGED=$1 RD_GN_STAGES=$2"
RD_GN_STAGED=$1 RD_GN_STAGES=$2 timeout 300 python tools/microbench.py gn 2>&1 | grep groupnorm | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', d['bench'], d['case'], round(d['ms'], 4), 'ms', round(d['frac_of_hbm'], 3))
"
done
for ST in 0 1 0 1; do
RD_GN_STAGED=$ST timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --gpu-reference off --stage-times --no-e2e > gpurun_out/bench_gn${ST}_$T.json 2> gpurun_out/bench_gn${ST}_$T.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_gn${ST}_$T.json').read().strip().splitlines()[-1])
print('staged', $ST, round(d['value'],2), d['stage_ms_last_step_rank0'], d['output_check']['ok'], d['gpu_launches'])
PY
done
