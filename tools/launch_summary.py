#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum [+ dram__bytes_read.sum, dram__bytes_write.sum]) of bench.py:
picks ONE steady-state forward (from the n-th stem kernel to the next), prints a per-launch table and per-class totals.

    launch_summary.py launches.csv [forward_index] [--csv out.csv] [--json out.json]
"""
import collections, csv, json, re, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1.0}


def load(fn):
    rows = list(csv.DictReader([l for l in open(fn) if not l.startswith("==")]))
    byid = collections.OrderedDict()
    for r in rows:
        d = byid.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    return list(byid.values())


def short(name):
    m = re.search(r"(\w+_kernel\w*)(<[^(]*>)?", name)
    s = (m.group(1) + (m.group(2) or "")) if m else name[:40]
    return s.replace("__nv_bfloat16", "bf16").replace("(bool)", "")


def klass(name):
    if "conv_tc" in name or "ffn_fused" in name or "bneck_tail" in name: return "gemm_tcgen05"
    if "enc_attn_fused" in name: return "attention"
    if "stem" in name: return "stem"
    if "attention" in name: return "attention"
    if "layernorm" in name or "cast_addpos" in name: return "norm"
    if "conv_simt" in name or "linear_simt" in name: return "gemm_cuda_core"
    return "other"


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    L = load(args[0])
    which = int(args[1]) if len(args) > 1 else 4
    stems = [i for i, d in enumerate(L) if "stem_tc_kernel" in d["name"] or "stem_kernel" in d["name"]]
    a, b = stems[which], stems[which + 1]
    fwd = L[a:b]
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    out_rows = []
    for i, d in enumerate(fwd):
        t = d.get("gpu__time_duration.sum", 0.0) / 1e3
        rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        c = tot[klass(d["name"])]
        c[0] += 1; c[1] += t; c[2] += rd; c[3] += wr
        out_rows.append((i, short(d["name"]), d["grid"], d["block"], round(t, 2), round(rd / 1e6, 2), round(wr / 1e6, 2)))
    if "--csv" in sys.argv:
        with open(sys.argv[sys.argv.index("--csv") + 1], "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["launch", "kernel", "grid", "block", "time_us", "dram_read_MB", "dram_write_MB"])
            w.writerows(out_rows)
    else:
        for r in out_rows:
            print(*r)
    summ = {k: {"launches": v[0], "time_us": round(v[1], 1), "dram_read_MB": round(v[2] / 1e6, 1), "dram_write_MB": round(v[3] / 1e6, 1)}
            for k, v in tot.items()}
    summ["total"] = {"launches": len(fwd), "time_us": round(sum(v[1] for v in tot.values()), 1),
                     "dram_read_MB": round(sum(v[2] for v in tot.values()) / 1e6, 1),
                     "dram_write_MB": round(sum(v[3] for v in tot.values()) / 1e6, 1)}
    print(json.dumps(summ, indent=1))
    if "--json" in sys.argv:
        json.dump(summ, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
