set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v4
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "coalign or bias_residual or groupnorm" > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
grep -E "passed|failed|rror|exit=|^FAILED" gpurun_out/pytest_gpu_$T.log | tail -15
for c in 1 2 5 6; do
RD_COALIGN_CFG=$c timeout 200 python tools/microbench.py align --iters 300 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d.get('bench')=='coalign': print('cfg $c', d['case'], round(d['us_per_iter'],2), 'us/iter', round(d['GBs'],0), 'GB/s', round(d['frac_of_hbm'],3), 'loss', round(d['lossN'],6))
"
done
timeout 300 python tools/microbench.py gn 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['bench'], d['case'], round(d['ms'],3), 'ms', round(d['GBs'],0), 'GB/s', round(d['frac_of_hbm'],3))
    except Exception: print(l.strip()[:200])
"
