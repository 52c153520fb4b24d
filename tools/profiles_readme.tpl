# profiles/ — measured evidence, named per round (all on B200, sm_100a, this pool). The round-1 section comes first; the CURRENT numbers are in "Round 2" further down.

| file | what | how |
|---|---|---|
| `r01_bench_stream_cfg2.json` | headline bench line of ROUND 1 (decode_stream_kernel; superseded by `r02_bench_default.json`, see "Round 2" below) incl. roofline, gemm_rooflines, cpu_baseline | `python bench.py` |
| `r01_bench_reference_arm.json` | reference arm: the reference's own CPU path (oracle/_ref, 16 host cores) | `python bench.py --impl reference --steps 3 --warmup 1` |
| `r01_bench_stream_cfg1_0p6b.json`, `r01_bench_stream_utt30.json` | the same bench on configs[0] (0.6B, 11 s) and one configs[4] unit (1.7B, 30 s, 128 tokens) | `python bench.py --workload cfg1|utt30` |
| `r01_bench_stream_cfg3.json`, `_cfg4.json`, `_cfg5.json` | configs[2] (0.6B, -S 20 over a 3600 s recording = 180 segments, 4 sequences per decode step), configs[3] (0.6B stream, 2 s chunks over 60 s, per-chunk latency), configs[4] (1.7B, 64 x 30 s utterances, 2 sequences per decode step) | `python bench.py --workload cfg3|cfg4|cfg5 [--utterances 64]` |
| `r01_stream_ncu_summary.json`, `r01_stream_ncu_full_raw.csv.gz` | `ncu --set full` capture of one decode_stream_kernel launch (16 greedy steps): DRAM traffic, throughput, pipes | command inside the JSON; raw page = `ncu -i … --page raw --csv` |
| `r01_gemm_tc_ncu_summary.json` | `ncu --set full` of the tcgen05 GEMM (128x256 tiles) at M=8192, K=4096, N=8192 | command inside the JSON |
| `r01_launches_stream_cfg2.csv.gz`, `.summary.md` | ncu launch list (device time of every kernel) of 2 utterances of the bench workload | `ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/profile_utt.py 1.7b 2` after the same command exited 0 without ncu |
| `r01_gemm_skinny_sweeps.txt` | skinny split-K GEMM at the bench shapes: ticket vs cluster/DSMEM reduction, dependent launch, fused hi/lo MMA, split-target sweeps, dropped variants | `tools/gemm_bench.py` under the environment switches of DESIGN §9 |
| `r01_prefill_ablation.txt` | in-graph marginal cost of every launch class of the prefill chain (GEMMs, attention, norms) | `QASR_PREFILL_ABLATE=<mask> python tools/profile_utt.py 1.7b 6` |
| `r01_stream_phase_breakdown.txt` | clock64 phase stamps of the decode kernel (CTA 0 / last CTA), per-unit trace | `tools/mega_prof.py`, `tools/mega_trace.py` |
| `r01_stream_microbench.txt` | cluster-16 feasibility, 2-D TMA box streaming rate, all-to-all exchange latency (vs protocol, replicas, CTA count), tuning sweeps, 2-GPU lines | `tools/microbench/*.cu`, `tools/sk_sweep.sh`, `tools/sk_variants.sh` |
| `r01_mega2_ncu_summary.json`, `r01_megakernel_phase_breakdown.txt`, `r01_bench_megakernel_v1.json`, `r01_bench_graph_decode.json`, `r01_launches_graph_decode_cfg2.csv.gz` | earlier decode paths of this round (per-phase kernels in a CUDA graph; grid-barrier megakernels), kept as the comparison | as named |

Regenerate: `python tools/update_profiles.py <prefix of the gpurun_out captures>` (this file comes from `tools/profiles_readme.tpl`).

## Headline (configs[1]: 1.7B, 3.64 s utterance, 32 tokens)

{b[value]:.1f} x realtime device-timed, {b[e2e][value]:.1f} x end to end through the C ABI with host buffers ({b[ms_per_step]:.1f} ms per utterance:
mel {b[stage_ms][mel_ms]:.2f} + encoder {b[stage_ms][enc_ms]:.2f} + prefill {b[stage_ms][prefill_ms]:.2f} + decode {b[stage_ms][decode_ms]:.2f} ms); the reference's CPU path on the box's 16 host cores:
{ref[value]:.2f} x realtime ({ref[ms_per_step]:.0f} ms). Greedy ids identical (`cpu_baseline.ids_match_gpu`). 2 GPUs 252 x, 4 GPUs 495 x, 8 GPUs 1037 x (weak scaling 0.99 of linear; measured when one GPU ran 131 x).

## Decode step, Qwen3-ASR-1.7B (3 458 793 472 algorithmic bytes per step at ~78 cached positions)

| path | ms / step | achieved GB/s | fraction of measured HBM peak (6542 GB/s) |
|---|---:|---:|---:|
| per-phase kernels in a CUDA graph (`QASR_DECODE=graph`) | 1.494 | 2316 | 0.354 |
| grid-barrier megakernel v1 (FFMA consumer, 1-D bulk rings) | 1.360 | 2544 | 0.389 |
| grid-barrier megakernel v2 (mma.sync consumer, 2-D TMA boxes; `QASR_DECODE=mega2`) | 1.326 | 2609 | 0.399 |
| **decode_stream_kernel** (default) | **{b[roofline][ms_per_launch]:.3f}** | **{b[roofline][achieved]:.0f}** | **{b[roofline][frac]:.3f}** ({nominal:.0f} % of the 8 TB/s nominal) |

ncu (`r01_stream_ncu_summary.json`): {ncu[dram_bytes_read]:,.0f} B of DRAM reads for 16 steps = {ncu[dram_bytes_per_step]:,.0f} B per step =
1.00 x the algorithmic bytes (nothing is re-read); `dram__bytes_read` {ncu[dram_read_TBps]:.2f} TB/s while the kernel runs;
tensor pipe (HMMA) {ncu[tensor_pipe_active_pct]:.0f} % active, shared-memory wavefronts {ncu[smem_wavefronts_pct_of_peak]:.0f} % of peak: neither bounds the kernel. What is
left is the chain of 5 all-to-all exchanges per layer (`r01_stream_phase_breakdown.txt`: ~20 us per layer
against 15.4 us of HBM time; an exchange through L2 costs 1.1-1.9 us on an idle B200,
`r01_stream_microbench.txt`).

Other workloads (same kernel): 30 s utterance, ~470 cached keys: {u30[roofline][ms_per_launch]:.3f} ms/step = {u30[roofline][frac]:.2f}; 0.6B (1 233 715 200 B per step):
{c1[roofline][ms_per_launch]:.3f} ms/step = {c1[roofline][achieved]:.0f} GB/s = {c1[roofline][frac]:.2f}: latency-bound (13 us per layer against 4.8 us of HBM time), was 1.0 ms with the
grid-barrier kernel.

## Several sequences per decode step (qasr_cuda_transcribe_batch)

| workload | sequences / step | ms / step | decoder tokens/s | realtime factor (1 GPU) | one sequence per step |
|---|---:|---:|---:|---:|---:|
| configs[2]: 0.6B, -S 20, 3600 s recording, 180 segments | 4 | {c3[roofline][ms_per_launch]:.3f} | {c3[decoder_tok_s]:.0f} | **{c3[value]:.0f} x** | 361 x |
| configs[4]: 1.7B, 64 x 30 s utterances, 128 tokens each | 2 | {c5[roofline][ms_per_launch]:.3f} | {c5[decoder_tok_s]:.0f} | **{c5[value]:.0f} x** | 256 x |
| configs[3]: 0.6B stream, 2 s chunks (replicas only) | 1 | {c4[roofline][ms_per_launch]:.3f} | {c4[decoder_tok_s]:.0f} | {c4[value]:.0f} x; chunk latency p50 {c4[chunk_latency_ms][p50]:.1f} ms, p95 {c4[chunk_latency_ms][p95]:.1f} ms | — |

(`roofline.frac` of these lines divides the bytes of ONE weight pass by the step time; B sequences share that pass.)

## tcgen05 GEMM path (encoder / prefill), `gemm_rooflines` in the bench line

| shape | time | achieved | of peak |
|---|---:|---:|---:|
| {g[0][shape]} (weight stream, skinny split-K kernel) | {g[0][us]:.1f} us | {g[0][achieved]:.0f} GB/s | {g[0][frac]:.2f} of the measured HBM peak |
| {g[1][shape]} | {g[1][us]:.0f} us | {g[1][achieved]:.0f} TFLOP/s (hi+lo MMAs) | {g[1][frac]:.2f} of the measured bf16 peak ({g[1][peak]:.0f}, burst) |
| {g[2][shape]} (128x256 tiles) | {g[2][us]:.0f} us | {g[2][achieved]:.0f} TFLOP/s (hi+lo MMAs) | {g[2][frac]:.2f} of the measured bf16 peak; ncu: {gemm_us:.0f} us, tensor pipe {gemm_pipe:.1f} % |

The f32-activation reference is reproduced by issuing MMA(A_hi, W) and MMA(A_lo, W) per k-block (one MMA over the
[hi | lo] planes in the skinny kernel): both are counted as tensor work; useful flops (2MNK) are half of that. At the single-utterance shapes of the bench (M = 47/61 rows) the GEMMs
are weight streams at 5-21 us each (split-K partials reduced across a thread-block cluster through distributed shared
memory; programmatic dependent launch lets the next GEMM's weight tiles stream while the previous kernels finish). The epilogue transposes every 32 x 32
accumulator block through shared memory so that stores are row-contiguous (thread-per-row stores bounded the medium-M
GEMMs: 30 s gate/up 80 -> 58 us, 8192 x 8192 x 4096 845 -> 676 us).

## Launch list of the bench workload (cold-cache, serialised: compare SHARES)

{launch}
`decode_stream_kernel` share of the serialised time vs `decode_ms / (mel+enc+prefill+decode)` in the un-profiled bench
line = {b[stage_ms][decode_ms]:.2f} / {b[ms_per_step]:.2f} = {dec_share:.1f} %: the shares agree. (`sk_retile_kernel` runs once at model load and is not
part of a step.)
