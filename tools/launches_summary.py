"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) of `bench.py --steps 2 --warmup 1`
into markdown: launches, total and mean time per kernel, and the share of each kernel in one full-histogram step.
usage: python tools/launches_summary.py gpurun_out/<name>_launches.csv profiles/<name>_bench_1gpu.json profiles/<name>_launches_summary.md"""
import collections, csv, json, re, sys

src, bench, out_path = sys.argv[1:4]
rows = []
with open(src) as f:
    lines = f.readlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
for r in csv.DictReader(lines[start:]):
    if r["Metric Name"] == "gpu__time_duration.sum":
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, float(r["Metric Value"]) / 1e6))
agg = collections.OrderedDict()
per = collections.OrderedDict()
for n, ms in rows:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ms
    per.setdefault(n, []).append(ms)
med = {n: sorted(v)[len(v) // 2] for n, v in per.items()}      # the bench's brute-force leg launches the sweep kernel with culling off
line = json.load(open(bench))
tag = re.search(r"(r\d\w)_", out_path).group(1)
out = ["# ncu launch list of `python bench.py --steps 2 --warmup 1` (%s)\n" % tag,
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c %d --csv` (the first %d launches; per-launch times are cold-cache" % (len(rows), len(rows)),
       "and serialised, so only the SHARES are comparable with the bench line).  Raw list: `%s_launches_bench_steps2.csv`;" % tag,
       "this file: `python tools/launches_summary.py`.\n",
       "| kernel | launches | total ms | mean ms per launch | median |", "|---|---|---|---|---|"]
for n, (c, t) in agg.items():
    out.append("| `%s` | %d | %.3f | %.4f | %.4f |" % (n, c, t, t / c, med[n]))
step = [("block_bbox_kernel", 1), ("pair_list_kernel", 2), ("scan2_kernel", 1), ("bin_table_kernel", 1), ("sweep_records_kernel", 1),
        ("void full_hist_warp_kernel<1, 0, 1>", 1), ("symmetrise_kernel", 1), ("epilogue_kernel", 1)]
tot = sum(med[n] * k for n, k in step if n in agg)
out += ["\n(medians: the brute-force leg of the bench launches the same sweep kernel with culling off, ~0.4 s per launch)",
        "\nOne full-histogram step (cfg5, 1 M atoms) = box pass + 2 x pair list + scan + bin table + sweep records + sweep + symmetrise + epilogue:\n",
        "| kernel | ms per step | share |", "|---|---|---|"]
for n, k in step:
    if n in agg:
        ms = med[n] * k
        out.append("| `%s` | %.4f | %.1f %% |" % (n, ms, 100 * ms / tot))
sweep = agg["void full_hist_warp_kernel<1, 0, 1>"]
rl = line["roofline"]
out.append("\nSum %.3f ms per step under ncu; the bench line (`%s_bench_1gpu.json`) has %.2f ms per step with `roofline.kernel_ms_per_launch` %.2f ms"
           % (tot, tag, line["ms_per_step"], rl["kernel_ms_per_launch"]))
out.append("(the sweep kernel = %.1f %% of the step there, %.1f %% here)." % (100 * rl["kernel_ms_per_launch"] / line["ms_per_step"],
                                                                          100 * med["void full_hist_warp_kernel<1, 0, 1>"] / tot))
open(out_path, "w").write("\n".join(out) + "\n")
print("wrote", out_path)
