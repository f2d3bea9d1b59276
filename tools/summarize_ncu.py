"""Summarise ncu artefacts from gpurun_out/ into small, tracked text files under profiles/.

    python tools/summarize_ncu.py <round-tag> <gpurun_out/dir>
"""
import collections, csv, json, os, subprocess, sys

tag, src = sys.argv[1], sys.argv[2]
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
os.makedirs(out, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "dram__bytes_read.sum.per_second"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    return rows[hdr], rows[hdr + 1], rows[hdr + 2:]


for rep in sorted(f for f in os.listdir(src) if f.endswith(".ncu-rep")):
    try:
        hdr, units, rows = raw(os.path.join(src, rep))
    except StopIteration:                                   # truncated / unreadable report: keep the previous summary
        print(f"skipped {rep}: ncu could not read it", file=sys.stderr)
        continue
    ki = hdr.index("Kernel Name")
    lines = [f"# {rep}: ncu --set full --clock-control none (cold-cache, serialised replays; use for shares and per-kernel metrics, not for bench values)"]
    for r in rows:
        lines.append(f"\n== {r[ki][:90]}")
        for h, u, v in zip(hdr, units, r):
            if h in KEYS and v != "":
                lines.append(f"   {h:80s} {v:>18s} {u}")
    open(os.path.join(out, f"{tag}_{rep.replace('.ncu-rep', '')}.txt"), "w").write("\n".join(lines) + "\n")

lc = os.path.join(src, "launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[h + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0][-60:]
        tot[name] += v; cnt[name] += 1
    s = sum(tot.values())
    with open(os.path.join(out, f"{tag}_launches.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: per-kernel device time over the captured launch window\n")
        f.write("# (cold-cache, serialised: compare SHARES with bench.py's roofline.kernel_share_of_step, not absolutes)\n")
        f.write("# command: bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-legs --games 37888, launches 1500..2100 of the campaign;\n")
        f.write("# the network's share grows with the campaign size (more searches per launch): ~82 % here, 97 % at the bench default\n")
        f.write("# of 303,104 games (roofline.kernel_share_of_step) -- ncu costs ~45 ms per intercepted launch, so the default size is\n")
        f.write("# out of reach for a launch list.  k_net_lat rows = launches that returned at once because the device-side batch\n")
        f.write("# was above the latency shape's range (since then only issued when the ply's searching-game count allows it).\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"{k:60s} launches={cnt[k]:5d} total_ms={v / 1e6:10.3f} avg_us={v / cnt[k] / 1e3:10.1f} share={100 * v / s:5.1f}%\n")
bj = os.path.join(src, "bench.json")
if os.path.exists(bj):
    d = json.load(open(bj))
    json.dump(d, open(os.path.join(out, f"{tag}_bench.json"), "w"), indent=1)
print("profiles written:", sorted(os.listdir(out)))
