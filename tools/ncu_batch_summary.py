"""Summarise an `ncu --set full --import-source on` capture of batch_kernel into markdown (profiles/*_batch_ncu_summary.md).
usage: python tools/ncu_batch_summary.py gpurun_out/<name>.ncu-rep profiles/<name>_summary.md
Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv --print-source cuda,sass`; the phase table maps
source lines of fullrmc_b200/csrc/store.cu (as checked out) to the phases of the kernel."""
import collections, csv, io, os, subprocess, sys

rep, out_path = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, vals = rows[0], rows[1], rows[2]
R = dict(zip(hdr, zip(units, vals)))
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
out = ["# ncu --set full: batch_kernel<ORTHO_FAST> (a run of up to 32 proposals resolved on the device), cfg5 (1M atoms)\n",
       "Command: `ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 4 -c 1 python tools/probe_batch.py cfg5 456`\n"
       "(the fifth batch launch of the run: 32 single-atom proposals, PDF + S(Q) models, hs 1000, nQ 400).  ncu flushes caches between\n"
       "replays and serialises, so the absolute duration below (cold 16 MB store from HBM, cold instruction cache) is larger than the\n"
       "~150 us a launch takes back to back in the benchmark; the instruction counts and their distribution are what this capture is for.\n"
       "Raw report: `%s` (scratch, not committed); this file: `python tools/ncu_batch_summary.py`.\n" % rep,
       "| metric | value | unit |\n|---|---|---|"]
for w in want:
    if w in R:
        out.append("| %s | %s | %s |" % (w, R[w][1], R[w][0]))
out.append("\n## Warp stall reasons (per issued instruction)\n| reason | stalled warps per issue |\n|---|---|")
for h in sorted(stalls, key=lambda h: -float(R[h][1] or 0)):
    out.append("| %s | %s |" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), R[h][1]))

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))
cur, hdr2, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        hdr2 = r
        isamp, iinst = hdr2.index("# Samples"), hdr2.index("Instructions Executed")
        continue
    if hdr2 and len(r) > isamp and r[0].isdigit():
        try:
            s = int(r[isamp]) if r[isamp] not in ("-", "") else 0
            i = int(r[iinst]) if r[iinst] not in ("-", "") else 0
        except ValueError:
            continue
        a = agg.setdefault((cur, int(r[0])), [0, 0, r[1]])
        a[0] += s
        a[1] += i
ts, ti = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
out.append("\n## Source lines with the most warp-stall samples (`--page source`, %d samples, %d warp instructions in all)\n"
           "| file:line | samples | %% | instructions | %% | source |\n|---|---|---|---|---|---|" % (ts, ti))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:22]:
    out.append("| %s:%d | %d | %.1f | %d | %.1f | `%s` |" % (k[0], k[1], v[0], 100 * v[0] / ts, v[1], 100 * v[1] / ti,
                                                            v[2].strip()[:80].replace("|", "\\|")))

src = open(os.path.join(root, "fullrmc_b200/csrc/store.cu")).read().split("\n")


def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]:
            return i + 1
    raise SystemExit("marker not found: " + pat)


k0 = find("batch_kernel(float4 *__restrict__ atoms")
delta = find("// ---- (2) delta pass of the whole batch")
rounds = find("// ---- (3) rounds")
decide = find("// chi2 of every slot: models in defer_mask")
commit = find("// ---- commit the accepted proposals (all CTAs)")
end = find("}  // namespace frmc", commit)
epi0 = find("__device__ __forceinline__ void epilogue_run")
epi_g = find("// ---- 1. r-space function")
epi_chi = find("{   // chi^2 summation schedule")
epi_sq = find("// ---- 3. S(Q) slice")
epi_end = find("epilogue_kernel(const ModelSet ms", epi_sq)
gw0, gw1 = find("__device__ __forceinline__ void grid_arrive"), find("__device__ __forceinline__ void resolve_body")
regions = [("epilogue: tables, setup", epi0, epi_g), ("epilogue: G(r) (counts -> r-space function)", epi_g, epi_chi),
           ("epilogue: chi2 of r-space models, schedule", epi_chi, epi_sq), ("epilogue: S(Q) rows + terms", epi_sq, epi_end),
           ("grid barrier (arrive / spin)", gw0, gw1), ("kernel head: clear, moved atoms, masks", k0, delta),
           ("delta pass (box, reach test, sweep, hits)", delta, rounds),
           ("round: plan builder, evaluation set-up, barrier call sites", rounds, decide),
           ("decide: chi2 sums, walk, next plan, outputs", decide, commit), ("commit + corrections + end", commit, end)]
tab = collections.OrderedDict()
for (f, l), (s, i, _) in agg.items():
    key = None
    if f == "store.cu":
        for name, a, b in regions:
            if a <= l < b:
                key = name
                break
    if key is None:
        key = "common.cuh (dist2, blocks_far, bin rule)" if f == "common.cuh" else "other (intrinsics headers, helpers of store.cu)"
    t = tab.setdefault(key, [0, 0])
    t[0] += s
    t[1] += i
out.append("\n## Where the launch spends its instructions and stall samples\n(one launch, 148 CTAs x 16 warps; samples are taken over ALL "
           "warps, so warps parked at a CTA barrier while thread 0 spins on the grid barrier weigh heavily)\n\n"
           "| phase | warp instructions | % | stall samples | % |\n|---|---|---|---|---|")
for name, (s, i) in sorted(tab.items(), key=lambda kv: -kv[1][1]):
    out.append("| %s | %d | %.1f | %d | %.1f |" % (name, i, 100 * i / ti, s, 100 * s / ts))
open(out_path, "w").write("\n".join(out) + "\n")
print("wrote", out_path)
