#!/bin/bash
mkdir -p gpurun_out
timeout 80 python tools/sched_probe.py --max-batch 128 --conc 128,256 --waves 4 2>&1 | tee gpurun_out/r2z_probe128.log
