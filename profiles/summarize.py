"""Turns the raw ncu exports a gpurun call leaves in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py <tag>        # e.g. r1_final: reads gpurun_out/{raw,sass,launches}_<tag>.csv

raw_<tag>.csv       ncu -i rep --page raw --csv                  (one `--set full` capture of k_sweep0<MEASURE>)
sass_<tag>.csv      ncu -i rep --page source --csv --print-source sass
launches_<tag>.csv  ncu --metrics gpu__time_duration.sum --clock-control none ... --csv  (launch list of bench.py)
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw_summary(tag):
    rows = list(csv.reader(open(os.path.join(OUT, f"raw_{tag}.csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for i, h in enumerate(hdr):
        if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            d[h] = {"unit": units[i], "value": vals[i]}
    return d


def phase_summary(tag):
    """Executed instructions and warp-state samples per code region between the kernel's barriers."""
    rows = list(csv.reader(open(os.path.join(OUT, f"sass_{tag}.csv"))))
    hdr = rows[1]
    iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[2:] if len(r) > iE]
    tot_s, tot_e = sum(int(r[iN]) for r in data), sum(int(r[iE]) for r in data)
    bars = [k for k, r in enumerate(data) if "BAR.SYNC" in r[iS]]
    edges = [0] + [b + 1 for b in bars] + [len(data)]
    # k_sweep0 as of r1_final: barrier after the mbarrier init, after the strip has landed, after the measurement,
    # after every half-sweep (folded into one region here: they sit inside inlined copies of the same function), ...
    names = ["setup + mbarrier init", "TMA issue, threshold table, wait for the strip (mbarrier)",
             "measure (level-0 correlators, block to level 1)",
             "half-sweeps (pass 1 + queue pass 2; the barriers between them are inside)", "(barrier tail of the last half-sweep)",
             "fence.proxy.async + TMA bulk store", "exit"]
    out = []
    for k, (a, b) in enumerate(zip(edges[:-1], edges[1:])):
        seg = data[a:b]
        s, e = sum(int(r[iN]) for r in seg), sum(int(r[iE]) for r in seg)
        st = {h[6:]: sum(int(r[i]) for r in seg) for i, h in stall}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
        mix = collections.Counter()
        for r in seg:
            op = r[iS].split()
            if not op:
                continue
            name = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
            mix[name.split(".")[0]] += int(r[iE])
        out.append({"region": names[k] if k < len(names) else f"region {k}", "sass_instructions": b - a,
                    "share_of_executed_instructions": round(e / tot_e, 4), "share_of_warp_samples": round(s / tot_s, 4),
                    "top_stalls_pct_of_region_samples": {h: round(100.0 * v / max(s, 1), 1) for h, v in top},
                    "top_opcodes_share_of_region": {o: round(c / max(e, 1), 3) for o, c in mix.most_common(8)}})
    return {"warp_samples": tot_s, "executed_warp_instructions": tot_e, "regions": out}


def launch_summary(tag):
    rows = [r for r in csv.reader(open(os.path.join(OUT, f"launches_{tag}.csv"))) if len(r) > 5]
    h = next(r for r in rows if r[0] == "ID")
    data = rows[rows.index(h) + 1:]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in data:
        v = float(r[iv].replace(",", ""))
        agg[r[ik]].append(v / 1000.0 if r[iu] == "ns" else v)
    tot = sum(sum(v) for v in agg.values())
    return [{"kernel": k, "launches": len(v), "avg_us": round(sum(v) / len(v), 2), "share_of_listed_time": round(sum(v) / tot, 4)}
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))]


def main():
    tag = sys.argv[1]
    res = {"tag": tag}
    if os.path.exists(os.path.join(OUT, f"raw_{tag}.csv")):
        res["ncu_set_full_k_sweep0_measure"] = raw_summary(tag)
    if os.path.exists(os.path.join(OUT, f"sass_{tag}.csv")):
        res["phases"] = phase_summary(tag)
    if os.path.exists(os.path.join(OUT, f"launches_{tag}.csv")):
        res["launch_list"] = launch_summary(tag)
    dst = os.path.join(ROOT, "profiles", f"ncu_summary_{tag}.json")
    with open(dst, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
