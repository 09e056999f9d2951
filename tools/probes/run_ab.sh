#!/bin/bash
B=tools/probes/bin
{
for b in xb_v6 xb_v7; do echo "== $b"; for fl in 0 1 2048; do timeout 10 $B/$b 1 5 27648 $fl 20; done; timeout 10 $B/$b 8 5 15552 0 10; timeout 10 $B/$b 8 10 3888 0 20; timeout 10 $B/$b 8 20 1008 0 20; timeout 10 $B/$b 3 5 130 0 5; done
echo "== v6 trace"; timeout 10 $B/xb_v6_trace 1 5 27648 0 10 | head -9
echo "== v6 trace midld"; timeout 10 $B/xb_v6_trace 1 5 27648 1 10 | head -9
echo "== v7 trace"; timeout 10 $B/xb_v7_trace 1 5 27648 0 10 | head -9
} > gpurun_out/xattn_ab.log 2>&1
cat gpurun_out/xattn_ab.log
