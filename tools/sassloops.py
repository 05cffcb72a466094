"""List the loops (backward branches) of one kernel in a cuobjdump -sass dump with their static instruction mix.
usage: python tools/sassloops.py <object or .so> <kernel-name-substring>"""
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
        print(name, len(lst), "instructions")
        addr = {a: i for i, (a, _) in enumerate(lst)}
        for i, (a, t) in enumerate(lst):
            m = re.search(r"BRA(?:\.\S+)*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
                lo = addr[int(m.group(1), 16)]
                body = [x[1] for x in lst[lo:i + 1]]
                ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", b).split()[0].split(".")[0] for b in body)
                print(f"  loop {lst[lo][0]:#x}..{a:#x}: {len(body)} instr  " + " ".join(f"{k}:{v}" for k, v in ops.most_common(18)))

if __name__ == "__main__":
    main()
