"""Per-kernel DRAM bytes and durations from `ncu --page raw --csv` exports -> profiles/<round>_traffic.json (read by bench.py).
usage: python tools/traffic_from_ncu.py <pairs_in_launch> rows=<raw.csv> fused=<raw.csv> > profiles/r2_traffic.json
The LAST launch of every kernel name in a file is taken (warm buffers)."""
import csv, json, re, sys


def unit_scale(u):
    u = u.strip().lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0,
            "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "s": 1e3}.get(u, 1.0)


def read(path):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(h)}
    out = {}
    for r in rows[2:]:
        if len(r) < len(h):
            continue
        name = re.sub(r"^(void )?(jtk::)?", "", r[ix["Kernel Name"]]).split("(")[0]
        g = lambda k: float(r[ix[k]].replace(",", "")) * unit_scale(units[ix[k]])
        out[name] = {"read": round(g("dram__bytes_read.sum") / 1e9, 6), "write": round(g("dram__bytes_write.sum") / 1e9, 6),
                     "ms": round(g("gpu__time_duration.sum"), 3)}
    return out


def main():
    pairs = int(sys.argv[1])
    res = {"pairs_in_launch": pairs}
    for a in sys.argv[2:]:
        variant, path = a.split("=", 1)
        k = read(path)
        seq = {n: v for n, v in k.items() if any(t in n for t in ("fwdrows", "bwdtable", "finalize", "modtable_fused"))}
        res[variant] = {"kernel": " + ".join(seq), "dram_bytes_read": sum(v["read"] for v in seq.values()) * 1e9,
                        "dram_bytes_write": sum(v["write"] for v in seq.values()) * 1e9, "per_kernel_gb": k,
                        "gpu_time_ms": round(sum(v["ms"] for v in seq.values()), 3),
                        "source": f"{path} (ncu --set full --clock-control none, tools/prof.py --reps 1, last launch of every kernel)"}
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
