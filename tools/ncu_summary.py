"""Summaries of ncu captures for profiles/ (run here, no GPU needed):
    python tools/ncu_summary.py raw <rep> <title>      key metrics of every launch in the report
    python tools/ncu_summary.py launches <csv> <title> per-kernel totals of a gpu__time_duration launch list
    python tools/ncu_summary.py stalls <rep> <title>   warp-stall sampling totals + hottest instructions
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return [r for r in csv.reader(io.StringIO(out)) if r]


def raw(rep, title):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    print(f"# {title}")
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print(f"# kernel: {name}")
        print("metric,value,unit")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"{m},{vals[i]},{units[i]}")


def launches(path, title):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    tot, agg = 0.0, {}
    for r in rows:
        name, unit, val = r[4], r[-2], float(r[-1].replace(",", ""))
        ms = val * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}[unit]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
        tot += ms
    print(f"# {title}")
    print("# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes")
    print(f"# total device time {tot:.2f} ms over {len(rows)} launches")
    print("kernel,launches,total_ms,share")
    for name, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
        print(f"\"{name[:140]}\",{n},{ms:.3f},{ms / tot:.4f}")


def stalls(rep, title):
    rows = ncu_csv(rep, "source")
    hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
    data = []  # first kernel of the report only
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) != len(hdr) or r == hdr:
            break
        data.append(r)
    ci = {h: i for i, h in enumerate(hdr)}
    cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ci["# Samples"]]) for r in data)
    print(f"# {title}")
    print(f"# warp-state sampling, {tot} samples over {len(data)} SASS instructions")
    print("stall_reason,samples,share")
    agg = {c: sum(int(r[ci[c]]) for r in data) for c in cols}
    for c, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
        print(f"{c},{v},{v / tot:.3f}")
    print("# hottest instructions: samples, times executed, SASS")
    for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:12]:
        print(f"# {r[ci['# Samples']]:>7} {r[ci['Instructions Executed']]:>10}  {r[ci['Source']].strip()[:90]}")


if __name__ == "__main__":
    {"raw": raw, "launches": launches, "stalls": stalls}[sys.argv[1]](sys.argv[2], sys.argv[3])
