set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v19
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "groupnorm or fused_modules or bias_residual or vae" > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
tail -3 gpurun_out/pytest_gpu_$T.log
timeout 300 python tools/microbench.py gn --full 2>&1 | tee gpurun_out/mb_gn_$T.log | python -c "
import sys, json
for l in sys.stdin:
    if not l.startswith('{'): print(l.rstrip()[:200]); continue
    d = json.loads(l)
    if 'no SiLU' in d['bench']: continue
    print('  ', d['bench'], d['case'], round(d['ms'], 4), 'ms', round(d['frac_of_hbm'], 3))
"
