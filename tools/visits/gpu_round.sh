#!/bin/bash
# One GPU-box visit: parity tests (each group in its own process so that a faulting kernel does not
# poison the rest), microbenchmarks, bench.py.  Everything is logged under gpurun_out/.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 6 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run t_lib 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "library" -p no:cacheprovider
run t_align 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "coalign or merge or min_and_prep or renorm" -p no:cacheprovider
run t_ddim 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ddim" -p no:cacheprovider
run t_f32 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fp32" -p no:cacheprovider
run t_xattn_tc 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tcgen05" -p no:cacheprovider
run t_xattn_rest 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "xattn and not tcgen05 and not fp32" -p no:cacheprovider
run t_proc 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "processor" -p no:cacheprovider
run t_pipe 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
if [ "${RD_SKIP_MB:-0}" != "1" ]; then
run mb_align 600 python tools/microbench.py align --iters 200
run mb_attn 600 python tools/microbench.py attn --flags 0 1
fi
for extra in "$@"; do run "extra_$(echo $extra | tr -c 'a-zA-Z0-9' '_' | cut -c1-40)" 900 bash -c "$extra"; done
cat gpurun_out/summary.txt
