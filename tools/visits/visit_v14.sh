set -u
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
T=v14
for ST in 0 1; do
echo "== RD_GN_STAGED=$ST"
RD_GN_STAGED=$ST timeout 300 python tools/microbench.py gn --full 2>&1 | grep "groupnorm+silu" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', d['case'], round(d['ms'], 4), 'ms', round(d['frac_of_hbm'], 3))
"
done
for ST in 0 1; do
RD_GN_STAGED=$ST RD_SB=16 RD_VAE_BS=36 timeout 400 python tools/profile_step.py fast 60 > gpurun_out/step_profile_gn${ST}_$T.txt 2>&1
grep -E "step wall|gn_|cutlass3x|xattn_fwd_kernel<__half, 0>|bias_residual|Self CUDA time" gpurun_out/step_profile_gn${ST}_$T.txt | cut -c1-100,200-330
done
