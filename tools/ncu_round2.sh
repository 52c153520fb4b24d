#!/bin/bash
# ncu --set full captures of the kernels VERDICT r1 asked evidence for (run under gpurun, one GPU).
# Usage: bash tools/ncu_round2.sh   -> gpurun_out/r2_ncu_*.ncu-rep
set -u
OUT=gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-strong --no-cpu-baseline"
P="python tools/batch_profile.py 1.7b 16 30 4 1"
$B > $OUT/r2_ncu_plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
$P > $OUT/r2_ncu_plain_batch.log 2>&1 || { echo "plain batch failed"; exit 1; }
NCU="ncu --set full --clock-control none --import-source on"
# single-utterance chains of the bench workload (1.7B, 3.64 s: encoder M = 47, prefill M = 61)
$NCU -k regex:gemm_tc_skinny_kernel -s 530 -c 8 -o $OUT/r2_ncu_skinny_prefill -f $B > /dev/null 2>&1
$NCU -k regex:attn_prefill_kernel -s 60 -c 2 -o $OUT/r2_ncu_attn_prefill -f $B > /dev/null 2>&1
$NCU -k regex:attn_windowed_kernel -s 60 -c 2 -o $OUT/r2_ncu_attn_windowed -f $B > /dev/null 2>&1
# batched path (16 x 30 s utterances, 1.7B)
$NCU -k regex:gemm_tc_kernel -s 204 -c 4 -o $OUT/r2_ncu_gemm_batched_prefill -f $P > /dev/null 2>&1
$NCU -k regex:gemm_tc_kernel -s 6 -c 4 -o $OUT/r2_ncu_gemm_batched_encoder -f $P > /dev/null 2>&1
$NCU -k regex:"attn_decode_batch_kernel|attn_prefill_batch_kernel|im2col_stage_kernel|conv1_kernel" -c 8 -o $OUT/r2_ncu_batched_misc -f $P > /dev/null 2>&1
$NCU -k regex:attn_decode_batch_kernel -s 30 -c 2 -o $OUT/r2_ncu_attn_decode_batch -f $P > /dev/null 2>&1
ls -la $OUT/*.ncu-rep
