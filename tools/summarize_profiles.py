"""Turn the ncu outputs of tools/profile_run.sh (gpurun_out/) into the committed summaries under profiles/."""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1d"

def launches(path, title):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1)
        short = re.sub(r"<.*", "", re.sub(r"\(.*", "", row["Kernel Name"]))
        a = agg.setdefault(short, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1; a[1] += v; tot += v; n += 1
    out = [f"## {title}\n\n{n} launches, {tot/1e6:.1f} ms of kernel time (cold-cache, serialised under ncu: compare SHARES).\n\n",
           "| kernel | launches | total ms | share | grid | block |\n|---|---:|---:|---:|---|---|\n"]
    for k, (c, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        out.append(f"| `{k}` | {c} | {t/1e6:.3f} | {100*t/tot:.2f}% | {g} | {b} |\n")
    return "".join(out), {k: v[1] / 1e6 for k, v in agg.items()}

def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    return dict(zip(r[0], zip(r[2], r[1])))

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_uniform.sum", "launch__shared_mem_per_block_dynamic"]

def to_bytes(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)

md = [f"# profiles/{tag}_launches.md -- ncu launch lists (tools/profile_run.sh)\n\n",
      "    ncu --metrics gpu__time_duration.sum --clock-control none -s <3 warm-up steps> -c <one step> --csv python bench.py --steps 1 --warmup 3 [--workload config1]\n\n"]
for name, title in (("launches_config2.csv", "bench workload: config 2 (270270 x 135135 after the transpose of tools/rank, p = 42013), one echelonize step"),
                    ("launches_config1.csv", "config 1 (20000 x 20000, 5 per row), one echelonize step")):
    p = os.path.join(G, name)
    if os.path.exists(p):
        text, _ = launches(p, title)
        md.append(text + "\n")
open(os.path.join(P, f"{tag}_launches.md"), "w").write("".join(md))

traffic = {}
md = [f"# profiles/{tag}_ncu_full.md -- `ncu --set full --clock-control none --import-source on`, one launch per kernel\n\n",
      "Reports (scratch, not committed): gpurun_out/prof_<kernel>.ncu-rep; the tables are what they contain.\n"]
names = {"k_greedy_ooo": ("greedy_pivot_search", "greedy cycle-free pivot search, out-of-order commits (config 2)"),
         "k_panel_solve_flow": ("panel_solve", "dataflow batched triangular solve (config 2; first batch of the step = the 100-row density estimate)"),
         "k_kahn_async": (None, "asynchronous Kahn levels of the pivot DAG (config 2)"),
         "k_rref_panel_cluster": (None, "dense echelon panel factorisation, 32 columns, cluster of 8 CTAs (config 2)"),
         "k_umma_gemm_packed": (None, "tcgen05 int8 limb-split modular product, block elimination of config 1 (1000 x <=6813 x 1000, 2 limbs, 128 x 64 tiles)"),
         "k_umma_gemm_packed_8192": ("umma_gemm_8192", "tcgen05 int8 limb-split modular product, 8192 x 8192 x 8192, 2 limbs, 128 x 128 tiles (tools/gemm_bench.py)")}
for k, (key, title) in names.items():
    rep = os.path.join(G, f"prof_{k}.ncu-rep")
    if not os.path.exists(rep):
        continue
    d = raw(rep)
    md.append(f"\n## `{k}` -- {title}\n\n| metric | value | unit |\n|---|---:|---|\n")
    for w in WANT:
        if w in d:
            md.append(f"| {w} | {d[w][0]} | {d[w][1]} |\n")
    stalls = sorted(((float(v[0].replace(",", "")), kk) for kk, v in d.items() if "issue_stalled" in kk and kk.endswith("_per_warp_active.pct")), reverse=True)[:6]
    md.append("\nTop stall reasons (% of warp-active cycles): " + ", ".join(f"{kk.split('stalled_')[1].split('_per')[0]} {v:.1f}" for v, kk in stalls) + "\n")
    if key and "dram__bytes_read.sum" in d:
        traffic[key + "_per_launch"] = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
open(os.path.join(P, f"{tag}_ncu_full.md"), "w").write("".join(md))
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
for f in (f"bench_{tag}.json", f"bench_{tag}_config1.json", f"bench_{tag}_config3.json", f"bench_{tag}_n2.json", f"bench_{tag}_reference.json", "gemm_bench_42013.txt", "gemm_bench_65537.txt", "gemm_bench_2147483629.txt"):
    src = os.path.join(G, f)
    if os.path.exists(src):
        open(os.path.join(P, f), "w").write(open(src).read())
print(open(os.path.join(P, f"{tag}_ncu_full.md")).read())
