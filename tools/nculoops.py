"""Dynamic view of an `ncu --page source --csv` export: for every loop (backward branch) of the first kernel, the executed warp
instructions inside its address range, per unit, with the opcode mix, shared-memory wavefronts and stall samples.
usage: python tools/nculoops.py <source.csv> <units> [min_share]"""
import csv, sys, re, collections


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
    ins = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name": break
        if len(r) < len(hdr) or r[0] == "Address": continue
        src = r[ix["Source"]].strip()
        op = re.sub(r"^@!?U?P\d+\s+", "", src)
        op = op.split()[0].split(".")[0] if op else "?"
        ins.append(dict(addr=int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]]), src=src, op=op,
                        n=int(r[ix["Instructions Executed"]] or 0), samp=int(r[ix["# Samples"]] or 0),
                        wave=int(r[ix["L1 Wavefronts Shared"]] or 0), ideal=int(r[ix["L1 Wavefronts Shared Ideal"]] or 0),
                        stalls={k[6:]: int(r[ix[k]] or 0) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}))
    pos = {d["addr"]: i for i, d in enumerate(ins)}
    tot = sum(d["n"] for d in ins)
    tsamp = sum(d["samp"] for d in ins)
    loops = []
    for i, d in enumerate(ins):
        m = re.search(r"BRA(?:\.\S+)*\s+(?:\S+,\s*)?0x([0-9a-f]+)", d["src"])
        if m:
            t = int(m.group(1), 16)
            if t in pos and pos[t] <= i:
                loops.append((pos[t], i))
    print(f"total {tot / units:.1f} instr per unit, {len(ins)} static, {len(loops)} loops")
    # innermost-first accounting: a loop's own share excludes nested loops that are listed
    loops.sort(key=lambda ab: ab[1] - ab[0])
    owned = [False] * len(ins)
    for lo, hi in loops:
        body = [k for k in range(lo, hi + 1) if not owned[k]]
        n = sum(ins[k]["n"] for k in body)
        if n < min_share * tot:
            continue
        for k in body: owned[k] = True
        ops = collections.Counter()
        for k in body: ops[ins[k]["op"]] += ins[k]["n"]
        wave = sum(ins[k]["wave"] for k in body); ideal = sum(ins[k]["ideal"] for k in body)
        samp = sum(ins[k]["samp"] for k in body)
        trip = max(ins[k]["n"] for k in body)
        print(f"loop {lo}..{hi} ({len(body)} own static): {n / units:.1f}/unit = {100.0 * n / tot:.1f}% instr, {100.0 * samp / max(tsamp, 1):.1f}% samples, "
              f"waves {wave / units:.1f} (ideal {ideal / units:.1f}), max exec {trip / units:.2f}/unit")
        print("    " + " ".join(f"{k}:{v / units:.1f}" for k, v in ops.most_common(16)))
        st = collections.Counter()
        for k in body:
            for a, b in ins[k]["stalls"].items(): st[a] += b
        tt = max(sum(st.values()), 1)
        print("    stalls: " + " ".join(f"{a}:{100.0 * b / tt:.0f}%" for a, b in st.most_common(7)))
        if len(sys.argv) > 4:
            hot = sorted(body, key=lambda k: -ins[k]["samp"])[:int(sys.argv[4])]
            for k in sorted(hot):
                print(f"      {ins[k]['samp']:6d}  {ins[k]['src'][:90]}")
    rest = [k for k in range(len(ins)) if not owned[k]]
    n = sum(ins[k]["n"] for k in rest)
    print(f"outside listed loops: {n / units:.1f}/unit = {100.0 * n / tot:.1f}%")


if __name__ == "__main__":
    main()
