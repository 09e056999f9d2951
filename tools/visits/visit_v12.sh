set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v12
( timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu_$T.log )
grep -E "cfg1:|fp16:|passed|failed|rror|exit=" gpurun_out/pytest_gpu_$T.log | tail -8
timeout 300 python tools/microbench.py gn > gpurun_out/mb_gn_$T.log 2>&1; cut -c1-220 gpurun_out/mb_gn_$T.log
timeout 300 python tools/microbench.py gn --only "dec L3" > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_apply|gn_stats" -s 6 -c 2 -f -o gpurun_out/prof_gn_$T python tools/microbench.py gn --only "dec L3" > gpurun_out/ncu_gn_$T.log 2>&1
echo "ncu gn exit=$?"
