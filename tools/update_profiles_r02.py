"""Rewrites the '# Round 2' section of profiles/README.md from the committed r02_* artefacts (python tools/update_profiles_r02.py)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
d = json.load(open(os.path.join(P, "r02_bench_default.json")))
d148 = json.load(open(os.path.join(P, "r02_bench_default_148cta.json")))  # 1-GPU line of the build the 2- / 8-GPU lines were measured on (148-CTA decode grid)
d2 = json.load(open(os.path.join(P, "r02_bench_default_2gpu.json")))
d8 = json.load(open(os.path.join(P, "r02_bench_default_8gpu.json")))
ref = json.load(open(os.path.join(P, "r02_bench_reference_arm.json")))
ncu = json.load(open(os.path.join(P, "r02_ncu_summary.json")))
st, st2, st8 = d["extra"]["strong"], d2["extra"]["strong"], d8["extra"]["strong"]
st148 = d148["extra"]["strong"]


def row(e):
    return (f"| `{e['kernel'].split('(')[0].replace('void ', '')}` | {int(e.get('grid', 0))} | {e.get('duration_us', 0):.1f} | {e.get('dram_bytes', 0) / 1e6:.1f} | "
            f"{e.get('dram_pct', 0):.1f} | {e.get('tensor_pipe_pct', 0):.1f} | {e.get('l2_throughput_pct', 0):.1f} | {e.get('warps_active_pct', 0):.1f} | {int(e.get('regs', 0))} |")


names = {
    "r2_ncu_skinny_prefill": "single-utterance prefill chain (1.7B, M = 61): `gemm_tc_skinny_kernel<64,6>` - WO, gate/up, down, QKV of two layers",
    "r2_ncu_skinny_encoder_final": "single-utterance encoder chain on the final build (1.7B encoder, T = 47): `gemm_tc_skinny_kernel<64,6>` - QKV, WO, fc1, fc2 of two layers (`ncu -k regex:gemm_tc_skinny_kernel -s 300 -c 8 python tools/profile_utt.py 1.7b 2`)",
    "r2_ncu_attn_prefill": "`attn_prefill_kernel` (P = 61)", "r2_ncu_attn_windowed": "`attn_windowed_kernel` (T = 47)",
    "r2_ncu_gemm_batched_prefill": "batched prefill (16 x 30 s, 1.7B, M = 6464), persistent one-CTA `gemm_tc_kernel<256,3>` (before the 2-CTA variant) - gate/up, down, QKV, WO",
    "r2_ncu_gemm_batched_encoder": "batched encoder (8 x 30 s per pass): `gemm_tc_kernel` - conv3, conv_out, QKV ...",
    "r2_ncu_batched_misc": "batched conv stem of round 1: `conv1_kernel`, `im2col_stage_kernel` (stage 2 hi / lo, stage 3 hi / lo) - the im2col launches are gone since the implicit-GEMM conv",
    "r2_ncu_attn_decode_batch": "`attn_decode_batch_kernel` (16 sequences, ~410 cached positions)"}
tables = []
for k, t in names.items():
    tables.append(f"\n{t}:\n\n| kernel | grid | us | DRAM MB | DRAM % | tensor pipe % | L2 % | warps active % | regs |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    tables += [row(e) for e in ncu[k][:8]]

s = f"""

# Round 2

| file | what | how |
|---|---|---|
| `r02_bench_default.json` | **default bench line of round 2**: configs[1] headline + `roofline_0p6b` + `extra.strong` (configs[4] 256 x 30 s and configs[2] 3600 s / 180 segments on the ranks of the run) + `gemm_rooflines` (algorithmic and issued) + cpu_baseline | `python bench.py` (about 50 s on the box) |
| `r02_bench_default_148cta.json`, `r02_bench_default_2gpu.json`, `r02_bench_default_8gpu.json` | the line on 1, 2 and 8 GPUs of the build before the decode grid change (148 CTAs) | `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps 10 --warmup 3` |
| `r02_bench_reference_arm.json` | reference arm: the reference's own CPU path (`oracle/_ref`, 16 host cores) | `python bench.py --impl reference --steps 3 --warmup 1` |
| `r02_ncu_summary.json` | `ncu --set full` captures of the kernels VERDICT r1 asked evidence for: skinny GEMMs of the prefill chain at M = 61, `attn_prefill_kernel`, `attn_windowed_kernel`, the persistent large-tile GEMM in the batched prefill / encoder, conv stem kernels, `attn_decode_batch_kernel` | `bash tools/ncu_round2.sh` (each profiled command first exited 0 without ncu), summarised by `tools/ncu_summary.py` |
| `r02_launches_batched_64x30s.csv.gz` | ncu launch list of the first version of the batched path: 64 x 30 s utterances, 1.7B, 3 decode steps | `ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/batch_profile.py 1.7b 64 30 3 1` |
| `r02_decode_latency.txt` | the 0.6B decode latency study: exchange microbenchmarks (aligned vs shared sectors, strides, 4-byte phase words, clusters / DSMEM, L2-atomic all-reduce), phase stamps of the ring kernel, the register-window experiment, the TMA producer / consumer kernel and its sweeps (ring depth, L2 prefetch distance, pacing, padding) | `tools/microbench/ll_exchange2.cu`, `cluster_exchange.cu`, `tools/decode_ab.py`, `tools/mega_prof_fine.py` |
| `r02_ncu_rounds_0p6b.json` | `ncu --set full` of `decode_rounds_kernel<2>` (0.6B, 11 s, one launch = 16 greedy steps): 1234.2 MB of DRAM traffic per step = 1.0004 x algorithmic | `ncu --set full --clock-control none -k regex:decode_rounds_kernel -s 1 -c 2 python tools/profile_utt.py 0.6b 3 11.0 48`, `tools/ncu_summary.py` |
| `r02_batched_path.txt` | stage timings of the batched path vs group size, GEMM microbenchmark at the batched shapes (zero vs hashed operands, f32-store vs pipeline epilogues, one-CTA-per-tile vs persistent vs 2-CTA kernel), dropped variants | `tools/batch_profile.py`, `tools/gemm_bench.py batch|decode` |

Regenerate this section: `python tools/update_profiles_r02.py`.

## Headline and the north-star target model

configs[1] (1.7B, 3.64 s, 32 tokens): {d['value']:.1f} x realtime device-timed, {d['e2e']['value']:.1f} x end to end ({d['ms_per_step']:.2f} ms per utterance: mel
{d['stage_ms']['mel_ms']:.2f} + encoder {d['stage_ms']['enc_ms']:.2f} + prefill {d['stage_ms']['prefill_ms']:.2f} + decode {d['stage_ms']['decode_ms']:.2f} ms), decode step {d['roofline']['ms_per_launch']:.3f} ms = {d['roofline']['achieved']:.0f} GB/s =
**{d['roofline']['frac']:.3f}** of the measured HBM peak; ids identical to the compiled reference (`cpu_baseline.ids_match_gpu` = {d['cpu_baseline']['ids_match_gpu']}); the reference's CPU path on the
box's 16 host cores: {ref['value']:.2f} x realtime ({ref['ms_per_step']:.0f} ms). `roofline_0p6b` (0.6B, 11 s, 48 tokens): {d['roofline_0p6b']['ms_per_launch']:.3f} ms per step = {d['roofline_0p6b']['achieved']:.0f} GB/s =
**{d['roofline_0p6b']['frac']:.3f}** on `decode_rounds_kernel` (TMA producer warp + 8 consumer warps; 0.390 / 0.483 ms on the ring kernel; `r02_decode_latency.txt`) - bound by the five dependent
exchanges per layer (DESIGN 8.1). `qasr_cuda_step_logits` runs the same kernels, so the logits-level parity
tests pin what the bench times. 8 GPUs: {d8['value']:.0f} x (one replica per GPU, weak scaling).

## The splits BASELINE.json names (`extra.strong`)

| job | 1 GPU | 2 GPUs | 8 GPUs | sequences per weight pass (1 / 2 / 8 GPUs) | decode-step roofline, 1 GPU |
|---|---:|---:|---:|---:|---:|
| configs[4]: 256 x 30 s utterances, 1.7B, 128 tokens each | **{st['configs[4]']['value']:.0f} x** ({st['configs[4]']['ms_per_step'] / 1e3:.2f} s for 7680 s of audio, {st['configs[4]']['decoder_tok_s']:.0f} tokens/s; round 1: 405 x) | {st2['configs[4]']['value']:.0f} x | {st8['configs[4]']['value']:.0f} x | {st['configs[4]']['sequences_per_decode_step']} / {st2['configs[4]']['sequences_per_decode_step']} / {st8['configs[4]']['sequences_per_decode_step']} | {st['configs[4]']['roofline']['frac']:.2f} of the HBM peak (weights once + f32 KV rows of every sequence) |
| configs[2]: 3600 s recording, -S 20 -W 3, 180 segments, 0.6B | **{st['configs[2]']['value']:.0f} x** ({st['configs[2]']['ms_per_step'] / 1e3:.2f} s, {st['configs[2]']['decoder_tok_s']:.0f} tokens/s; round 1: 770 x) | {st2['configs[2]']['value']:.0f} x | {st8['configs[2]']['value']:.0f} x | {st['configs[2]']['sequences_per_decode_step']} / {st2['configs[2]']['sequences_per_decode_step']} / {st8['configs[2]']['sequences_per_decode_step']} | {st['configs[2]']['roofline']['frac']:.2f} |

2 GPUs: {st2['configs[4]']['value'] / st148['configs[4]']['value']:.2f} x / {st2['configs[2]']['value'] / st148['configs[2]']['value']:.2f} x of one GPU; 8 GPUs: {st8['configs[4]']['value'] / st148['configs[4]']['value']:.2f} x / {st8['configs[2]']['value'] / st148['configs[2]']['value']:.2f} x; configs[1] replicas (weak): {d2['value']:.0f} x and {d8['value']:.0f} x = {d8['value'] / d148['value'] / 8:.2f} of linear at 8.
(The 2- and 8-GPU lines were measured before the decode grid went from 148 to 132 CTAs; ratios are against the 1-GPU line of that build, `r02_bench_default_148cta.json`: {d148['value']:.1f} x, configs[4] {st148['configs[4]']['value']:.0f} x, configs[2] {st148['configs[2]']['value']:.0f} x. The grid only changes the single-sequence decode kernel, i.e. the configs[1] replicas.)
Strong scaling is bounded by the shard size, not by a collective (there is none): at 8 GPUs a rank holds 32 utterances / 22-23 segments, so
32 / 23 sequences share each pass over the weights instead of 128 / 90 (one GPU, 1.7B: 24.2k / 17.5k / 12.2k tokens/s at 128 / 64 / 32 sequences per
step) and the encoder / prefill GEMMs run at a quarter of the rows; slowest / mean rank {st8['configs[4]']['rank_ms']['imbalance']:.3f} (configs[4]) and {st8['configs[2]']['rank_ms']['imbalance']:.3f} (configs[2]).

## GEMM rooflines (`gemm_rooflines`: hashed non-zero operands, the epilogue of the pipeline)

| shape | us | 2MNK TFLOP/s | frac_algorithmic | frac_issued (hi + lo MMAs) |
|---|---:|---:|---:|---:|
"""
for g in d["gemm_rooflines"]:
    if g["bound"] == "tensor":
        s += f"| {g['shape']} | {g['us']:.1f} | {g['achieved']:.0f} | {g['frac_algorithmic']:.2f} | {g['frac_issued']:.2f} |\n"
    else:
        s += f"| {g['shape']} (weight stream) | {g['us']:.1f} | {g['achieved']:.0f} GB/s | {g['frac']:.2f} of the HBM peak | - |\n"
s += """
Peak = 1653.5 TFLOP/s (MEASURED_PEAKS.json, cuBLAS burst). The hi/lo split of the f32 activations issues two MMAs per k-block,
so `frac_algorithmic` = `frac_issued` / 2 by construction; the M = 25856 shapes run as CTA pairs (`tcgen05.mma.cta_group::2`). ncu on the
one-CTA persistent kernel showed the tensor pipe 74-92 % active on the batched prefill GEMMs. All-zero operands (round 1's
microbenchmark) read ~25 % higher than hashed ones (power), and the f32-store epilogue hid the cost of the SwiGLU / GELU / residual
epilogues: see `r02_batched_path.txt`.

## ncu --set full (`r02_ncu_summary.json`)
""" + "\n".join(tables) + """

Reading: the single-utterance chains are launch / latency bound (DRAM 8-22 %, 80-96 CTAs, 14-28 us per GEMM; attention 10-14 us
on 32-64 CTAs). The batched GEMMs are tensor-pipe bound (74-92 %). `attn_decode_batch_kernel` streams the f32 KV rows (2.1 TB/s
at 16 sequences / 128 CTAs, 4.6-5.3 TB/s at 64-128 sequences).
"""

# ---- ncu launch list of the bench workload (2 utterances through the C ABI): shares must agree with the stage times of the bench line
def launch_list_section(path):
    import collections, csv, gzip, re
    if not os.path.exists(path):
        return ""
    rows = list(csv.reader(l for l in gzip.open(path, "rt") if l.startswith('"')))
    ki, vi = rows[0].index("Kernel Name"), rows[0].index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        agg[n][0] += 1
        agg[n][1] += v / 1e3
    load_only = sum(v[1] for k, v in agg.items() if "sk_retile" in k)  # model load, not part of a step
    tot = sum(v[1] for v in agg.values()) - load_only
    t = ("\n## Launch list of the bench workload, round 2 (`r02_launches_default.csv.gz`; cold-cache, serialised: compare SHARES)\n\n"
         "`ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/profile_utt.py 1.7b 2` after the same command exited 0 without ncu: "
         f"2 utterances = {tot / 1e3:.2f} ms of serialised kernel time (without the load-time `sk_retile_kernel`, {load_only / 1e3:.2f} ms).\n\n"
         "| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    dec = 0.0
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if "sk_retile" in k or v[1] / tot < 0.0005:
            continue
        if "decode_rounds_kernel" in k or "decode_stream_kernel" in k:
            dec += v[1]
        t += f"| `{k}` | {v[0]} | {v[1]:.0f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f} % |\n"
    sm = d["stage_ms"]
    step = sm["mel_ms"] + sm["enc_ms"] + sm["prefill_ms"] + sm["decode_ms"]
    first = d["roofline"]["ms_per_launch"]  # the first greedy step runs inside the decode kernel but is booked under prefill_ms
    t += (f"\nThe decode kernel's share of the serialised time is {100 * dec / tot:.1f} %; in the un-profiled bench line its launches take "
          f"decode_ms + the first greedy step (booked under prefill_ms) = {sm['decode_ms']:.2f} + {first:.2f} of {step:.2f} ms = "
          f"{100 * (sm['decode_ms'] + first) / step:.1f} %: the shares agree (the cold-cache, serialised list inflates the short encoder / prefill kernels).\n")
    return t


s += launch_list_section(os.path.join(P, "r02_launches_default.csv.gz"))
readme = open(os.path.join(P, "README.md")).read()
i = readme.find("\n\n# Round 2")
open(os.path.join(P, "README.md"), "w").write((readme[:i] if i >= 0 else readme) + s)
print("profiles/README.md: round-2 section rewritten")
