#!/bin/bash
B=tools/probes/bin
{
for b in xb_v5_slow xb_v5_fast xb_v6_slow xb_v6_fast; do echo "== $b"; timeout 60 $B/$b 1 5 27648 0 20; timeout 60 $B/$b 8 5 15552 0 10; done
echo "== v6 trace"; timeout 60 $B/xattn_bench_trace 1 5 27648 0 10 | head -12
} > gpurun_out/xattn_ab.log 2>&1
cat gpurun_out/xattn_ab.log
