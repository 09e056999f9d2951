#!/bin/bash
B=tools/probes/bin
{
for fl in 0 2048 1024; do timeout 10 $B/xb_v7 1 5 27648 $fl 20; done
timeout 10 $B/xb_v7 8 5 15552 0 10
timeout 10 $B/xb_v7 8 10 3888 0 20
timeout 10 $B/xb_v7 8 20 1008 0 20
timeout 10 $B/xb_v7 8 20 252 0 20
timeout 10 $B/xb_v7 2 5 1000 0 20
timeout 10 $B/xb_v7 3 5 77 0 20
timeout 10 $B/xb_v7 3 5 130 0 20
timeout 10 $B/xb_v7 1 5 27648 0 20 1
timeout 10 $B/xb_v7_trace 1 5 27648 0 10
} > gpurun_out/xattn_v7.log 2>&1
cat gpurun_out/xattn_v7.log
