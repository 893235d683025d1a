"""Turn the ncu outputs of tools/r2_profile.sh (gpurun_out/) into the committed summaries under profiles/:

    profiles/<tag>_launches.md   per kernel family: launches, device time, share of the step, DRAM bytes (one step)
    profiles/<tag>_ncu_full.md   `ncu --set full` tables of the dominant kernels, with their stall reasons
    profiles/traffic.json        {workload: {kernel family: DRAM bytes per step}}  -- read by bench.py (roofline.traffic)

    python tools/summarize_profiles.py [tag]        (default r2)
"""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
UNIT = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
# kernel families of bench.py's roofline table
FAMILY = {"greedy_pivot_search": ("k_win_", "k_greedy", "k_longest_row", "k_strip_tags"),
          "panel_solve": ("k_panel_solve",)}


def launches(path, title, reps):
    """the csv holds `reps` echelonizations; the LAST one is the steady-state step"""
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = collections.OrderedDict()
    for row in csv.DictReader(lines):
        rid = row["ID"]
        r = rows.setdefault(rid, {"name": re.sub(r"<.*", "", re.sub(r"\(.*", "", row["Kernel Name"])).replace("void ", "").replace("sb::", ""),
                                  "grid": row["Grid Size"], "block": row["Block Size"], "ns": 0.0, "dram": 0.0})
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Name"] == "gpu__time_duration.sum":
            r["ns"] = v * UNIT.get(row["Metric Unit"], 1)
        elif row["Metric Name"].startswith("dram__bytes"):
            r["dram"] += v * BYTES.get(row["Metric Unit"], 1)
    seq = list(rows.values())
    step = seq[len(seq) - len(seq) // reps:] if reps > 1 else seq
    agg, tot = collections.OrderedDict(), 0.0
    for r in step:
        a = agg.setdefault(r["name"], [0, 0.0, 0.0, r["grid"], r["block"]])
        a[0] += 1; a[1] += r["ns"]; a[2] += r["dram"]; tot += r["ns"]
    out = [f"## {title}\n\n{len(step)} launches, {tot/1e6:.1f} ms of kernel time, {sum(a[2] for a in agg.values())/1e9:.2f} GB of DRAM traffic "
           "(cold-cache, serialised under ncu: compare SHARES).\n\n",
           "| kernel | launches | total ms | share | DRAM MB | grid | block |\n|---|---:|---:|---:|---:|---|---|\n"]
    for k, (c, t, d, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        out.append(f"| `{k}` | {c} | {t/1e6:.3f} | {100*t/tot:.2f}% | {d/1e6:.1f} | {g} | {b} |\n")
    fam = {}
    for f, prefixes in FAMILY.items():
        fam[f] = sum(a[2] for k, a in agg.items() if k.startswith(prefixes))
    return "".join(out), fam


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    return dict(zip(r[0], zip(r[2], r[1])))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]


def stalls(d):
    """warp stall reasons of the WarpStateStats section: cycles a warp spends stalled on <reason> per issued instruction"""
    found = []
    for k, (v, unit) in d.items():
        m = re.match(r"smsp__average_warps?_(?:latency_)?issue_stalled_(.+?)_per_(?:warp_active|issue_active)\.(pct|ratio)$", k)
        if m:
            try:
                found.append((float(v.replace(",", "")), m.group(1), m.group(2)))
            except ValueError:
                pass
    best = {}
    for v, name, kind in found:
        if name not in best or kind == "ratio":
            best[name] = (v, kind)
    return sorted(((v, n, kd) for n, (v, kd) in best.items()), reverse=True)[:7]


def main():
    os.makedirs(P, exist_ok=True)
    md = [f"# profiles/{tag}_launches.md -- ncu launch lists with DRAM bytes (tools/r2_profile.sh)\n\n",
          "    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv python tools/gpu_full.py <config>   (REPS=2: the second echelonization is the step below)\n\n"]
    traffic = {}
    for w, wl, title in (("c2", "config2", "bench workload: config 2 (270270 x 135135 after the transpose of tools/rank, p = 42013), one echelonize step"),
                         ("c1", "config1", "config 1 (20000 x 20000, 5 per row), one echelonize step"),
                         ("c3", "config3", "config 3 (171375 x 47271, 11 per row, --dense-threshold 0.01), one echelonize step")):
        path = os.path.join(G, f"{tag}_launches_{w}.csv")
        if os.path.exists(path):
            text, fam = launches(path, title, 2)
            md.append(text + "\n")
            traffic[wl] = fam
    open(os.path.join(P, f"{tag}_launches.md"), "w").write("".join(md))
    if traffic:
        json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)

    md = [f"# profiles/{tag}_ncu_full.md -- `ncu --set full --clock-control none --import-source on`, one launch per kernel\n\n",
          "Reports (scratch, not committed): gpurun_out/<tag>_prof_<kernel>.ncu-rep; the tables are what they contain.\n"]
    names = {"k_win_bfs": "greedy search, phase A: one snapshot search per window row (config 2, window 40 of ~75)",
             "k_win_resolve": "greedy search, phase B: ordered resolution of a window by inheritance of closed searches (config 2, window 40)",
             "k_panel_solve_flow2": "dataflow batched triangular solve, second version (config 2: the merged pass of the step)",
             "k_rref_panel_cluster": "dense echelon panel factorisation, 32 columns, cluster of 8 CTAs (config 2)",
             "k_umma_gemm_packed": "tcgen05 int8 limb-split modular product, block elimination of config 1 (1000 x <=6813 x 1000, 2 limbs)"}
    for k, title in names.items():
        rep = os.path.join(G, f"{tag}_prof_{k}.ncu-rep")
        if not os.path.exists(rep):
            continue
        d = raw(rep)
        md.append(f"\n## `{k}` -- {title}\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w in WANT:
            if w in d:
                md.append(f"| {w} | {d[w][0]} | {d[w][1]} |\n")
        st = stalls(d)
        if st:
            md.append("\nTop stall reasons (warps stalled on the reason per issue-active cycle, WarpStateStats): "
                      + ", ".join(f"{n} {v:.2f}" for v, n, kd in st) + "\n")
        else:
            md.append("\nStall reasons: not present in this report (keys with `stall`: "
                      + ", ".join(sorted(kk for kk in d if "stall" in kk)[:6]) + ")\n")
    open(os.path.join(P, f"{tag}_ncu_full.md"), "w").write("".join(md))
    for f in os.listdir(G):
        if f.startswith(f"bench_{tag}") and f.endswith(".json") or f in ("int8_peak.json",) or f.startswith("sanitizer_"):
            open(os.path.join(P, f), "w").write(open(os.path.join(G, f)).read())
    print(open(os.path.join(P, f"{tag}_ncu_full.md")).read())


if __name__ == "__main__":
    main()
