"""Turns the ncu exports of tools/gpu_ncu.sh / tools/gpu_launches.sh (gpurun_out/) into the markdown summary that is
committed under profiles/.

    python tools/ncu_summarize.py gpurun_out profiles/r01_ncu_summary_v5.md "<title line>"
"""
import collections
import csv
import json
import os
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def short_name(full):
    head = full.split("(")[0]
    base = head.split("::")[-1]
    if "conv_tc_kernel" in head:
        return "conv_tc_kernel<128>" if "<128>" in head or "(int)128" in head else "conv_tc_kernel<64>"
    if "ccl_merge_kernel" in head:
        return "ccl_merge_kernel<26>" if "26" in head else "ccl_merge_kernel<6>"
    return base


def launch_table(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(short_name(r[ix["Kernel Name"]]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[ix["Metric Value"]]) * UNIT[r[ix["Metric Unit"]]] * 1e3
    total = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total ms | share | avg ms |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.3f | %.1f%% | %.4f |" % (k, v[0], v[1], 100 * v[1] / total, v[1] / v[0]))
    conv = sum(v[1] for k, v in agg.items() if k.startswith("conv_tc"))
    return out, conv / total, total


def raw_table(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    return hdr, units, data, ix


CONV_METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
                "launch__shared_mem_per_block_dynamic"]


def main():
    src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
    md = ["# " + title, ""]
    lp = os.path.join(src, "launches.csv")
    if os.path.exists(lp):
        tab, share, total = launch_table(lp)
        md += ["## Launch list: device time per kernel of one `bench.py --steps 1 --warmup 0` run under "
               "`ncu --metrics gpu__time_duration.sum --clock-control none` (this repo's kernels only; three passes of the "
               "path over a 300-slice volume; cold-cache and serialised: compare SHARES)", ""] + tab + [""]
        md += ["conv_tc_kernel share of all kernel time under ncu: **%.1f%%** (%.1f ms of kernel time in total)." % (100 * share, total), ""]
    bp = os.path.join(src, "bench_final.json")
    if os.path.exists(bp):
        b = json.load(open(bp))
        r = b["roofline"]
        live = r["launches_timed"] / b["steps"] * r["avg_launch_ms"]
        md += ["bench.py's live CUDA-event figure (not under a profiler): %d conv launches x %.4f ms = %.1f ms of a %.1f ms step = "
               "**%.1f%%**; %.1f TFLOP/s algorithmic = %.3f of the measured bf16 peak." % (
                   r["launches_timed"] / b["steps"], r["avg_launch_ms"], live, b["ms_per_step"], 100 * live / b["ms_per_step"],
                   r["achieved"], r["frac"]), ""]
    cp = os.path.join(src, "prof_conv_raw.csv")
    if os.path.exists(cp):
        hdr, units, data, ix = raw_table(cp)
        md += ["## `ncu --set full -k regex:conv_tc_kernel -s 24 -c 4` (launches 25-28 of the bench: second 37-slice wave, "
               "down2.block0 .. down3.block3)", "", "| metric | " + " | ".join("launch %d" % (i + 1) for i in range(len(data))) + " |",
               "|---|" + "---|" * len(data)]
        for m in CONV_METRICS:
            if m in ix:
                md.append("| %s (%s) | %s |" % (m, units[ix[m]], " | ".join(d[ix[m]] for d in data)))
        tr = [(float(d[ix["dram__bytes_read.sum"]]) * UNIT[units[ix["dram__bytes_read.sum"]]] +
               float(d[ix["dram__bytes_write.sum"]]) * UNIT[units[ix["dram__bytes_write.sum"]]]) for d in data]
        md += ["", "Mean DRAM traffic of these launches: **%.1f Mbyte per launch** (bench.py `roofline.traffic`)." % (sum(tr) / len(tr) / 1e6), ""]
    sp = os.path.join(src, "prof_small_raw.csv")
    if os.path.exists(sp):
        hdr, units, data, ix = raw_table(sp)
        t, rd, wr = "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"
        dp, smp = "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"
        agg = collections.OrderedDict()
        for d in data:
            a = agg.setdefault(short_name(d[ix["Kernel Name"]]), [0, 0.0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += float(d[ix[t]]) * UNIT[units[ix[t]]]
            a[2] += float(d[ix[rd]]) * UNIT[units[ix[rd]]] + float(d[ix[wr]]) * UNIT[units[ix[wr]]]
            a[3] = max(a[3], float(d[ix[dp]]))
            a[4] = max(a[4], float(d[ix[smp]]))
        md += ["## `ncu --set full` over the non-convolution kernels (tools/profile_small_kernels.py: 74-slice 320x320 volume of a "
               "random-weight model, 37-slice waves)", "",
               "| kernel | launches captured | total ms | DRAM bytes (read+write) | achieved DRAM GB/s | max DRAM % of peak | max SM throughput % |",
               "|---|---|---|---|---|---|---|"]
        for k, a in agg.items():
            md.append("| %s | %d | %.3f | %.1f MB | %.0f | %.1f | %.1f |" % (k, a[0], a[1] * 1e3, a[2] / 1e6, a[2] / a[1] / 1e9 if a[1] else 0, a[3], a[4]))
        md.append("")
    open(dst, "w").write("\n".join(md) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
