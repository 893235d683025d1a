"""Markdown tables of DESIGN.md 5.4 / 6 from the bench snapshots under profiles/ (bench_r2*.json)."""
import json, os, sys
P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
def lines(name):
    path = os.path.join(P, name)
    return [json.loads(l) for l in open(path) if l.startswith("{")] if os.path.exists(path) else []
head = lines("bench_r2.json")[-1]
ref = lines("bench_r2_reference.json")[-1]
print("| workload | value (s, A resident in HBM) | e2e (s, host buffers) | pivots ms (greedy) | solve ms | dense ms (tensor GEMM) | dense mod-p Tfieldop/s |")
print("|---|---:|---:|---:|---:|---:|---:|")
for d in lines("bench_r2_all_workloads.json"):
    ph = d["phases_ms"]
    print(f"| {d['config']['workload'].split(':')[0]} | {d['value']:.4f} | {d['e2e']['value']:.4f} | {ph['ms_pivots']:.1f} ({ph['ms_pivots_greedy']:.1f}) | {ph['ms_solve']:.1f} | "
          f"{ph['ms_dense']:.1f} ({ph['ms_dense_gemm']:.1f}) | {d.get('dense_modp_tops', 0):.2f} |")
print()
print(f"headline (config 2, {head['steps']} steps): value {head['value']:.4f} s, e2e {head['e2e']['value']:.4f} s, reference arm {ref['value']:.3f} s on {ref['cpu_baseline']['cores']} cores "
      f"-> e2e ratio {ref['value'] / head['e2e']['value']:.1f}x; roofline {head['roofline']['kernel']} {head['roofline']['achieved']:.1f} GB/s = {100 * head['roofline']['frac']:.2f} % of {head['roofline']['peak']} GB/s; "
      f"panel_solve {head['kernels']['panel_solve']['algorithmic_GBps']:.0f} GB/s = {100 * head['kernels']['panel_solve']['frac_of_hbm_peak']:.1f} %; first call {head['first_call_s']:.2f} s; launches/step {head['gpu_launches'] / head['steps']:.0f}")
if "schur" in head and "rows_per_s" in head["schur"]:
    s = head["schur"]
    print(f"schur leg: {s['rows_per_s']:.0f} rows/s (GPU, C ABI, host arrays) vs {s.get('reference_rows_per_s', 0):.0f} rows/s (reference, {s.get('reference_cores')} cores)")
if "ingest" in head and "gpu_s" in head["ingest"]:
    g = head["ingest"]
    print(f"ingest leg: host loop {1e3 * g['host_s']:.1f} ms ({g['host_MBps']:.0f} MB/s), GPU {1e3 * g['gpu_s']:.1f} ms ({g['gpu_MBps']:.0f} MB/s)")
print()
print("| N | config 2 value (s) | e2e (s) | scale leg: config 4 echelonize + rref + kernel (s) | rref (s) | speed-up vs 1 GPU | NCCL bytes / step |")
print("|---:|---:|---:|---:|---:|---:|---:|")
for n in (2, 4, 8):
    for d in lines(f"bench_r2_n{n}.json")[-1:]:
        leg = d.get("scale_leg", {})
        print(f"| {n} | {d['value']:.4f} | {d['e2e']['value']:.4f} | {leg.get('value', 0):.2f} | {leg.get('rref_s', 0):.2f} | {leg.get('speedup_vs_single_gpu', 0):.2f} (single {leg.get('single_gpu_value', 0):.2f} s) | {leg.get('nccl_bytes_per_step', 0):.3g} |")
