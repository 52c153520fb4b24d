"""Summarise .ncu-rep captures (ncu -i ... --page raw --csv) into one JSON: python tools/ncu_summary.py out.json rep1 rep2 ..."""
import csv, io, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_insts",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_throughput_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "sm__cycles_active.avg": "sm_cycles_active",
}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        e = {"kernel": d.get("Kernel Name", "")[:110]}
        for k, name in KEYS.items():
            if k in d and d[k] != "":
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                u = units[hdr.index(k)]
                if k == "gpu__time_duration.sum":
                    v = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)
                if u in ("Kbyte",):
                    v *= 1e3
                if u in ("Mbyte",):
                    v *= 1e6
                if u in ("Gbyte",):
                    v *= 1e9
                e[name] = v
        if "dram_read_bytes" in e:
            e["dram_bytes"] = e.get("dram_read_bytes", 0) + e.get("dram_write_bytes", 0)
            if e.get("duration_us"):
                e["dram_gbs"] = e["dram_bytes"] / e["duration_us"] / 1e3
        res.append(e)
    return res


if __name__ == "__main__":
    out = {}
    for rep in sys.argv[2:]:
        out[rep.split("/")[-1].replace(".ncu-rep", "")] = load(rep)
    json.dump(out, open(sys.argv[1], "w"), indent=1)
    for name, ks in out.items():
        print("==", name)
        for e in ks:
            print("  %-60s grid %-8s %8.1f us  dram %6.1f%% (%7.0f GB/s, %6.1f MB)  tensor %5.1f%%  L2 %5.1f%%  warps %5.1f%%  regs %s" % (
                e["kernel"][:60], int(e.get("grid", 0)), e.get("duration_us", 0), e.get("dram_pct", 0), e.get("dram_gbs", 0), e.get("dram_bytes", 0) / 1e6,
                e.get("tensor_pipe_pct", 0), e.get("l2_throughput_pct", 0), e.get("warps_active_pct", 0), int(e.get("regs", 0))))
