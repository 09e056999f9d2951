set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v18
( timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "groupnorm or fused_modules" > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
tail -3 gpurun_out/pytest_gpu_$T.log
for OV in 2 8 24; do
echo "== RD_GN_OVERHEAD=$OV"
RD_GN_OVERHEAD=$OV timeout 300 python tools/microbench.py gn --full 2>&1 | grep "groupnorm+silu" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', d['case'], round(d['ms'], 4), 'ms', round(d['frac_of_hbm'], 3))
"
done
cp rollingdepth_b200/librd_b200.so /tmp/new.so
for which in new prev new prev; do
if [ $which = prev ]; then cp tools/probes/librd_b200_prev.so rollingdepth_b200/librd_b200.so; else cp /tmp/new.so rollingdepth_b200/librd_b200.so; fi
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --gpu-reference off --stage-times --no-e2e > gpurun_out/bench_${which}_$T.json 2> gpurun_out/bench_${which}_$T.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${which}_$T.json').read().strip().splitlines()[-1])
print('$which', round(d['value'],2), d['stage_ms_last_step_rank0'], d['output_check']['ok'], d['clocks'])
PY
done
cp /tmp/new.so rollingdepth_b200/librd_b200.so
