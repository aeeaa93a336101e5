#!/bin/bash
# round-2 GPU pass F: ncu evidence for every tensor-core kernel + launch lists (1 GPU; never wrap multi-rank commands)
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# prefill 4096: launch list (all kernels of the prefill) and --set full of one layer's GEMMs + attention
CL_GRAPH=0 timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:'gemm_tcgen05|attn_prefill|rmsnorm|rope_scatter|silu|embed_rows|batch_' -c 400 --csv \
   --log-file gpurun_out/r2f_prefill_launches.csv python tools/step_profile.py 1 4096 > gpurun_out/r2f_ncu_a.log 2>&1
CL_GRAPH=0 timeout 900 $NCU --set full --import-source on -k regex:'gemm_tcgen05|attn_prefill' -s 10 -c 5 -o gpurun_out/r2f_prefill -f \
   python tools/step_profile.py 1 4096 > gpurun_out/r2f_ncu_b.log 2>&1
# batched step B = 8 (ctx 1024): launch list + --set full of one layer's kernels
CL_GRAPH=0 timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:'gemm_tcgen05|attn_decode|batch_|step_|embed_kernel|gemv' -c 400 --csv \
   --log-file gpurun_out/r2f_batch8_launches.csv python tools/batch_profile.py 8 2 > gpurun_out/r2f_ncu_c.log 2>&1
CL_GRAPH=0 timeout 900 $NCU --set full --import-source on -k regex:'gemm_tcgen05_kernel<32|attn_decode_tc|batch_' -s 40 -c 10 -o gpurun_out/r2f_batch8 -f \
   python tools/batch_profile.py 8 1 > gpurun_out/r2f_ncu_d.log 2>&1
# persistent batched kernel (opt-in) and the single-sequence persistent kernel
CL_BATCH_MEGA=1 CL_GRAPH=0 timeout 600 $NCU --set full --import-source on -k regex:decode_mega_batch -s 1 -c 1 -o gpurun_out/r2f_mega_batch -f \
   python tools/batch_profile.py 8 2 > gpurun_out/r2f_ncu_e.log 2>&1
CL_GRAPH=0 timeout 600 $NCU --set full --import-source on -k regex:'decode_mega_kernel' -s 1 -c 1 -o gpurun_out/r2f_mega -f \
   python tools/step_profile.py 3 > gpurun_out/r2f_ncu_f.log 2>&1
ls -la gpurun_out/r2f_* | head -20
