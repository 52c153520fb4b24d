"""Regenerate profiles/ from a gpurun_out capture set: python tools/update_profiles.py <prefix> [launch_prefix] [prof_prefix]
(<prefix>_bench.log, _cfg1.log, _utt30.log, _cfg3.log, _cfg4.log, _cfg5.log, _stream_full.ncu-rep in gpurun_out/)."""
import collections, csv, gzip, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
pre = sys.argv[1]
lpre = sys.argv[2] if len(sys.argv) > 2 else None
ppre = sys.argv[3] if len(sys.argv) > 3 else None

def line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])

for src, dst in (("bench", "cfg2"), ("cfg1", "cfg1_0p6b"), ("utt30", "utt30"), ("cfg3", "cfg3"), ("cfg4", "cfg4"), ("cfg5", "cfg5")):
    f = os.path.join(G, f"{pre}_{src}.log")
    if os.path.exists(f):
        line(f)
        shutil.copy(f, os.path.join(P, f"r01_bench_stream_{dst}.json"))
rep = os.path.join(G, f"{pre}_stream_full.ncu-rep")
if os.path.exists(rep):
    raw = os.path.join(G, f"{pre}_stream_full_raw.csv")
    open(raw, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    rows = list(csv.reader(open(raw)))
    d = {h: v for h, v in zip(rows[0], rows[2])}
    g = lambda k: float(d[k].replace(",", ""))
    steps = 16
    out = {"kernel": d["Kernel Name"],
           "capture": "ncu --set full --clock-control none --import-source on -k regex:decode_stream --launch-skip 1 -c 1 python tools/profile_utt.py 1.7b 1 (one launch = 16 greedy steps, Qwen3-ASR-1.7B, 3.64 s utterance)",
           "steps_in_launch": steps, "gpu_time_ms": g("gpu__time_duration.sum"),
           "dram_bytes_read": g("dram__bytes_read.sum") * 1e9, "dram_bytes_write": g("dram__bytes_write.sum") * 1e6,
           "dram_bytes_per_step": (g("dram__bytes_read.sum") * 1e9 + g("dram__bytes_write.sum") * 1e6) / steps,
           "algorithmic_bytes_per_step": 3458793472, "dram_read_TBps": g("dram__bytes_read.sum.per_second"),
           "dram_read_pct_of_ncu_peak": g("dram__bytes_read.sum.pct_of_peak_sustained_elapsed"),
           "lts_sector_hit_rate_pct": g("lts__t_sector_hit_rate.pct"),
           "tensor_pipe_active_pct": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
           "smem_wavefronts_pct_of_peak": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
           "registers_per_thread": int(g("launch__registers_per_thread")), "grid": 148, "block": 512,
           "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active")}
    json.dump(out, open(os.path.join(P, "r01_stream_ncu_summary.json"), "w"), indent=1)
    with open(raw, "rb") as fi, gzip.open(os.path.join(P, "r01_stream_ncu_full_raw.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)
    print("ncu:", out["gpu_time_ms"] / steps, "ms/step", out["dram_bytes_per_step"], "B/step")
if lpre:
    src = os.path.join(G, f"{lpre}_launches.csv")
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    kn, mv = rows[hi].index("Kernel Name"), rows[hi].index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        try:
            v = float(r[mv].replace(",", ""))
        except (ValueError, IndexError):
            continue
        name = r[kn].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        lines.append(f"| `{k}` | {v[0]} | {v[1] / 1e3:.0f} | {v[1] / v[0] / 1e3:.1f} | {100 * v[1] / tot:.1f} % |")
    open(os.path.join(P, "r01_launches_stream_cfg2.summary.md"), "w").write(f"2 utterances = {tot / 1e6:.2f} ms of serialised kernel time.\n\n" + "\n".join(lines) + "\n")
    with open(src, "rb") as fi, gzip.open(os.path.join(P, "r01_launches_stream_cfg2.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)
    print("\n".join(lines[:5]))
if ppre:
    with open(os.path.join(P, "r01_stream_phase_breakdown.txt"), "w") as fo:
        fo.write("# decode_stream_kernel<1,5>, clock64 stamps of thread 0 of CTA 0 (always an attention CTA) and of the last CTA (tools/mega_prof.py)\n## Qwen3-ASR-1.7B\n")
        fo.write(open(os.path.join(G, f"{ppre}_prof17.log")).read())
        fo.write("## Qwen3-ASR-0.6B\n" + open(os.path.join(G, f"{ppre}_prof06.log")).read())
        fo.write("\n# per-unit trace, warp 0 of CTA 147, first layers (tools/mega_trace.py 147 1.7b): data is (almost) always already in the ring\n")
        fo.write("".join(open(os.path.join(G, f"{ppre}_trace147.log")).readlines()[:45]))


def write_readme():
    b, c1, u30 = line(os.path.join(P, "r01_bench_stream_cfg2.json")), line(os.path.join(P, "r01_bench_stream_cfg1_0p6b.json")), line(os.path.join(P, "r01_bench_stream_utt30.json"))
    c3, c4, c5 = (line(os.path.join(P, f"r01_bench_stream_cfg{i}.json")) for i in (3, 4, 5))
    ref = line(os.path.join(P, "r01_bench_reference_arm.json"))
    ncu = json.load(open(os.path.join(P, "r01_stream_ncu_summary.json")))
    gn = json.load(open(os.path.join(P, "r01_gemm_tc_ncu_summary.json")))["metrics"]
    launch = open(os.path.join(P, "r01_launches_stream_cfg2.summary.md")).read()
    g = b["gemm_rooflines"]
    tpl = open(os.path.join(ROOT, "tools", "profiles_readme.tpl")).read()
    txt = tpl.format(b=b, c1=c1, u30=u30, c3=c3, c4=c4, c5=c5, ref=ref, ncu=ncu, launch=launch, g=g,
                     gemm_pipe=float(gn["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]), gemm_us=float(gn["gpu__time_duration.sum"]),
                     dec_share=100 * b["stage_ms"]["decode_ms"] / b["ms_per_step"], nominal=100 * b["roofline"]["achieved"] / 8000)
    open(os.path.join(P, "README.md"), "w").write(txt)


write_readme()
