set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v5
export RD_COALIGN_CFG=5
CMD3="python tools/microbench.py align --only 1024x576 --iters 20"
timeout 300 $CMD3 > gpurun_out/plain3_$T.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:coalign_pass -s 8 -c 1 -f -o gpurun_out/prof_coalign_$T $CMD3 > gpurun_out/ncu_coalign_$T.log 2>&1
echo "ncu coalign exit=$?"
tail -3 gpurun_out/plain3_$T.log | cut -c1-300
