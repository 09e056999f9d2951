set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v20
( RD_GN_APPLY_SHORT=1 timeout 70 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "groupnorm or fused_modules" > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
tail -3 gpurun_out/pytest_gpu_$T.log
for M in 0 1; do
echo "== RD_GN_APPLY_SHORT=$M"
RD_GN_APPLY_SHORT=$M timeout 45 python tools/microbench.py gn --full 2>&1 | grep "groupnorm+silu" | tee -a gpurun_out/mb_gn_short${M}_$T.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', d['case'], round(d['ms'], 4), 'ms', round(d['frac_of_hbm'], 3))
"
done
