"""Summarise an `ncu --page source --csv` export of one kernel: executed warp instructions and stall samples by opcode,
and the hottest address ranges.  usage: python tools/ncusrc.py <source.csv> [n_pairs]"""
import csv, sys, re, collections

def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    ops = collections.Counter(); samp = collections.Counter(); stall = collections.Counter()
    tot = 0; wave = 0; wave_ideal = 0
    for r in rows[2:]:
        if r and r[0] == "Kernel Name": break  # first launch only
        if len(r) < len(hdr) or r[0] == "Address": continue
        src = r[ix["Source"]].strip()
        src = re.sub(r"^@!?U?P\d+\s+", "", src)
        op = src.split()[0].split(".")[0] if src else "?"
        n = int(r[ix["Instructions Executed"]] or 0)
        ops[op] += n; tot += n
        samp[op] += int(r[ix["# Samples"]] or 0)
        wave += int(r[ix["L1 Wavefronts Shared"]] or 0); wave_ideal += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                stall[k] += int(r[ix[k]] or 0)
    div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    print(f"total warp instructions {tot}  ({tot / div:.1f} per unit); shared wavefronts {wave} (ideal {wave_ideal}) = {wave / div:.1f} per unit")
    ts = sum(samp.values())
    for op, n in ops.most_common(28):
        print(f"  {op:10s} {n / div:12.2f}  {100.0 * n / tot:5.1f}%   samples {100.0 * samp[op] / max(ts, 1):5.1f}%")
    tt = sum(stall.values())
    print("stalls:", " ".join(f"{k[6:]}:{100.0 * v / tt:.1f}%" for k, v in stall.most_common(10)))

if __name__ == "__main__":
    main()
