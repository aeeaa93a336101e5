#!/bin/bash
# round-2 GPU pass A: parity (all -m gpu tests), bench line, cooperative-launch A/B, launch list + ncu full of the persistent kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
rm -f gpurun_out/parity_longctx.jsonl
( time timeout 1700 python -m pytest tests -m gpu -q -s -x -p no:cacheprovider ) > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 900 python bench.py --steps 64 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
CL_MEGA_COOP=1 timeout 400 python bench.py --steps 64 --warmup 5 --no-box --no-cpu-baseline --no-extra-configs > gpurun_out/r2a_bench_coop.json 2> gpurun_out/r2a_bench_coop.err
CL_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'decode_mega|gemv|step_|embed_kernel' -c 40 --csv \
   --log-file gpurun_out/r2a_launches.csv python tools/step_profile.py 3 > gpurun_out/r2a_ncu1.log 2>&1
CL_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 1 -c 1 -o gpurun_out/r2a_mega -f \
   python tools/step_profile.py 3 > gpurun_out/r2a_ncu2.log 2>&1
tail -5 gpurun_out/r2a_tests.log; tail -c 600 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err; tail -c 300 gpurun_out/r2a_bench_coop.json
