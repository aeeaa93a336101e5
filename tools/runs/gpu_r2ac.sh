#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/batch_step_profile.py 128 256 2>&1 | tail -16 | tee gpurun_out/r2ac_step_b128.log
timeout 200 python tools/batch_step_profile.py 64 256 2>&1 | tail -16 | tee gpurun_out/r2ac_step_b64.log
