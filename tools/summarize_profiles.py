#!/usr/bin/env python
"""Turn gpurun_out/ artefacts (ncu launch list, ncu --set full report, bench JSON, microbenchmarks) into the small
text summaries committed under profiles/.  Usage: python tools/summarize_profiles.py r01"""
import collections, csv, json, os, subprocess, sys
csv.field_size_limit(10**9)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
os.makedirs(P, exist_ok=True)

# ---- launch list
rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    n = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui] == "ns" else v
    agg.setdefault(n, []).append(v)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(P, "%s_ncu_launches.txt" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 : python bench.py --steps 2 --warmup 3 --no-cpu --dtype bf16 --no-graph --lanes 1\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write("%-52s %5s %10s %7s\n" % ("kernel", "n", "avg_us", "share"))
    for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write("%-52s %5d %10.1f %6.1f%%\n" % (n[:52], len(v), sum(v) / len(v), 100 * sum(v) / tot))

# ---- full capture
rep = os.path.join(G, "prof_conv.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
with open(os.path.join(P, "%s_ncu_conv_full.txt" % tag), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:conv_ -s 15 -c 5 : python bench.py --steps 2 --warmup 3 --no-cpu --dtype bf16 --no-graph --lanes 1\n")
    f.write("# one step of the regression trunk at batch 256: conv1 (conv_s2d_kernel), conv2..4 (conv_halo_kernel), conv5 + FC (conv_tma_kernel)\n")
    for r in rows[2:]:
        f.write("\n== %s\n" % r[H.index("Kernel Name")].split("(")[0])
        for w in want:
            if w in H:
                f.write("   %-68s %s %s\n" % (w, r[H.index(w)], rows[1][H.index(w)]))

# ---- alignment kernel
rep = os.path.join(G, "prof_align.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    with open(os.path.join(P, "%s_ncu_align.txt" % tag), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:align_tile -c 1 : python tools/bench_kernels.py align\n")
        f.write("# config C4: 4096 faces from 64 1080p frames -> 112x112x3 (algorithmic bytes 547 MB); times under ncu are cold-cache\n")
        for r in rows[2:]:
            f.write("\n== %s\n" % r[H.index("Kernel Name")].split("(")[0])
            for w in want + ["sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                             "smsp__sass_inst_executed_op_shared_ld.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"]:
                if w in H:
                    f.write("   %-68s %s %s\n" % (w, r[H.index(w)], rows[1][H.index(w)]))
            stalls = sorted(((float(r[i].replace(",", "")), h) for i, h in enumerate(H) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and
                             not h.endswith("_not_issued") and r[i]), reverse=True)[:8]
            f.write("   warp-stall samples: " + ", ".join("%s %d" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), v) for v, h in stalls) + "\n")

# ---- bench lines and microbenchmarks
for name in ("bench_bf16", "bench_bf16x3", "bench_fp32", "bench_reference", "bench_2gpu", "bench_4gpu", "bench_8gpu"):
    src = os.path.join(G, name + ".json")
    if os.path.exists(src):
        line = open(src).read().strip().splitlines()[-1]
        json.loads(line)
        open(os.path.join(P, "%s_%s.json" % (tag, name)), "w").write(line + "\n")
for src, dst in (("kernels.jsonl", "kernels.jsonl"), ("sweep_bf16.jsonl", "sweep_bf16.jsonl")):
    if os.path.exists(os.path.join(G, src)):
        open(os.path.join(P, "%s_%s" % (tag, dst)), "w").write(open(os.path.join(G, src)).read())
print("wrote", sorted(os.listdir(P)))
