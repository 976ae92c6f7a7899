#!/usr/bin/env python
"""profiles/: launch-list summary, top-kernel metrics and DRAM traffic from the ncu artefacts of one version.
usage: summarize_profiles.py <tag, e.g. r01_v5> <gpurun_out/xxx.ncu-rep>"""
import collections, csv, io, json, subprocess, sys
tag, rep = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(f"profiles/{tag}_launches.csv") if l.startswith('"')))
hdr = rows[0]; i_name = hdr.index("Kernel Name"); i_val = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[i_name].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += float(r[i_val].replace(",", ""))
tot = sum(v[1] for v in agg.values())
out = [f"# {tag} — ncu launch list summary, config 2, `python bench.py --steps 2 --warmup 3 --no-parity --no-config3` (device leg, kernel "
       "profile pass and the two end-to-end legs: k_slim_pack belongs to the slim end-to-end leg only)",
       "# per-launch times are cold-cache/serialised: compare SHARES", "kernel,launches,total_us,share"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k},{c},{t / 1000:.1f},{t / tot:.3f}")
open(f"profiles/{tag}_launch_summary.csv", "w").write("\n".join(out) + "\n")
print("\n".join(out[:9]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
summ = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    summ.setdefault(name, {k: r[idx[k]] for k in keys if k in idx})
json.dump(summ, open(f"profiles/{tag}_ncu_top_kernels.json", "w"), indent=1)
def num(v):
    return float(str(v).replace(",", ""))


t = {k: int((num(v["dram__bytes_read.sum"]) + num(v["dram__bytes_write.sum"])) * 1e6) for k, v in summ.items()}
# issue-slot roofline: warp instructions per second against what the SM sub-partitions can issue (1 per clock each)
SMSP, CLOCK_HZ = 148 * 4, 1.965e9
issue = {}
for k, v in summ.items():
    sec = num(v["gpu__time_duration.sum"]) * 1e-6  # ncu reports usecond here
    inst = num(v["smsp__inst_executed.sum"])
    issue[k] = {"warp_inst": inst, "ms": sec * 1e3, "achieved_ginst_per_s": inst / sec / 1e9, "peak_ginst_per_s": SMSP * CLOCK_HZ / 1e9,
                "frac": inst / sec / (SMSP * CLOCK_HZ), "issue_active_pct": num(v["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                "active_lanes_per_inst": num(v["smsp__thread_inst_executed_per_inst_executed.ratio"]),
                "occupancy_pct": num(v["sm__warps_active.avg.pct_of_peak_sustained_active"])}
json.dump({"source": f"profiles/{tag}_ncu_top_kernels.json (ncu --set full, config 2, one launch each; units as ncu reports them: Mbyte)",
           "per_launch_dram_bytes": t, "main_kernels_sum": sum(t.values()),
           "issue_roofline": {"note": "warp instructions issued per second vs 148 SMs x 4 sub-partitions x 1 instruction per clock at 1.965 GHz; "
                                      "per launch under ncu (cold caches, serialised), so read the fractions, not the times", "kernels": issue}},
          open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(summ, indent=1)); print(t)
