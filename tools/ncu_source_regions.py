"""Summarise `ncu --page source --csv` output: instructions executed, average active lanes and stall samples per
contiguous SASS region (regions split at the given opcode markers or every N instructions).

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_regions.py src.csv [chunk]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    tot_inst = sum(float(r[col["Instructions Executed"]]) for r in data)
    tot_samp = sum(float(r[col["# Samples"]]) for r in data)
    print("total warp instructions %.4g, samples %d, SASS lines %d" % (tot_inst, tot_samp, len(data)))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for a in range(0, len(data), chunk):
        seg = data[a:a + chunk]
        inst = sum(float(r[col["Instructions Executed"]]) for r in seg)
        thr = sum(float(r[col["Thread Instructions Executed"]]) for r in seg)
        samp = sum(float(r[col["# Samples"]]) for r in seg)
        st = {h: sum(float(r[col[h]]) for r in seg) for h in stalls}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        ops = {}
        for r in seg:
            op = r[col["Source"]].split()[0 if not r[col["Source"]].strip().startswith("@") else 1].split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        keyops = ",".join("%s%d" % (k, v) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
        if inst / max(tot_inst, 1) < 0.002 and samp / max(tot_samp, 1) < 0.002:
            continue
        print("%5d-%5d inst %5.1f%% lanes %4.1f samples %5.1f%%  %s   [%s]" % (
            a, a + len(seg), 100 * inst / tot_inst, thr / max(inst, 1), 100 * samp / max(tot_samp, 1),
            " ".join("%s=%.0f%%" % (k.replace("stall_", ""), 100 * v / max(samp, 1)) for k, v in top), keyops))


if __name__ == "__main__":
    main()
