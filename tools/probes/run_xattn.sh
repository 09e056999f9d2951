#!/bin/bash
# usage (on the GPU box): tools/probes/run_xattn.sh [binary]  -> gpurun_out/xattn_probe.log
B=tools/probes/bin/${1:-xattn_bench}
{
for fl in 0 2 2048; do timeout 10 $B 1 5 27648 $fl 20; done
timeout 10 $B 8 5 15552 0 10
timeout 10 $B 8 10 3888 0 20
timeout 10 $B 8 20 1008 0 20
timeout 10 $B 8 20 252 0 20
timeout 10 $B 2 5 1000 0 20
timeout 10 $B 3 5 77 0 20
timeout 10 $B 3 5 130 0 20
timeout 10 $B 1 5 27648 0 20 1
} > gpurun_out/xattn_probe.log 2>&1
cat gpurun_out/xattn_probe.log
