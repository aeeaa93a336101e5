#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/step_vs_b.py 256 2>&1 | tee gpurun_out/r2n_step_vs_b.log
timeout 600 python tools/sched_probe.py --conc 64 --jitter 0,64 2>&1 | tee gpurun_out/r2n_sched_probe2.log
CL_GRAPH=0 timeout 600 python tools/sched_probe.py --conc 64 --jitter 64 2>&1 | tee gpurun_out/r2n_sched_probe3.log
