"""Static register-move census of one kernel: instructions, FFMA2, MOV + IMAD.MOV in the whole kernel and in its loops.
usage: python tools/sassmov.py <object or .so> <kernel-name-substring>"""
import re, subprocess, sys, collections

def main():
    obj, pat = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, ins = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); ins[cur] = []; continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur:
            ins[cur].append((int(m.group(1), 16), m.group(2).strip()))
    for name, lst in ins.items():
        if pat not in name:
            continue
        def census(body):
            ops = collections.Counter()
            for b in body:
                b = re.sub(r"^@!?U?P\d+\s+", "", b)
                op = b.split()[0]
                ops["MOVES" if op in ("MOV", "IMAD.MOV.U32", "IMAD.MOV") else op.split(".")[0]] += 1
            return ops
        allops = census([t for _, t in lst])
        print(name, len(lst), "instr; moves", allops["MOVES"], "FFMA2", allops["FFMA2"], "LDS", allops["LDS"])
        addr = {a: i for i, (a, _) in enumerate(lst)}
        for i, (a, t) in enumerate(lst):
            m = re.search(r"BRA(?:\.\S+)*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
                lo = addr[int(m.group(1), 16)]
                ops = census([x[1] for x in lst[lo:i + 1]])
                if ops["FFMA2"] >= 40:
                    print(f"  loop {lst[lo][0]:#x}..{a:#x}: {i + 1 - lo} instr  moves {ops['MOVES']}  FFMA2 {ops['FFMA2']}  LDS {ops['LDS']}  FSEL {ops['FSEL']}  IMAD {ops['IMAD']}")

if __name__ == "__main__":
    main()
