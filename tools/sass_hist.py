"""SASS instruction histogram of a kernel's loops (no GPU needed).

    python tools/sass_hist.py <object-or-.so> <mangled-name-substring> [--dump]

Disassembles the matching function with cuobjdump, finds every backward branch (a loop), and prints for each loop body its
length and opcode mix, plus the HMMA count so that "instructions per super-tile" (= body / (#HMMA / 4)) can be read off.
Used to steer the GEMV decode loop towards the issue-slot budget in DESIGN.md and to commit the histogram under profiles/.
"""
import collections
import re
import subprocess
import sys


def disasm(obj, pat):
    names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.findall(r"Function : (\S+)", names)
    cand = [f for f in funcs if all(p in f for p in pat.split(","))]
    if not cand:
        raise SystemExit(f"no function matching {pat!r}; have e.g. {funcs[:5]}")
    fn = cand[0]
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return fn, ins


def opcode(text):
    t = re.sub(r"^@!?U?P\d\s+", "", text)
    return t.split()[0].split(".")[0] if t else "?"


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    fn, ins = disasm(obj, pat)
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    print(f"{fn}: {len(ins)} instructions")
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(?:!?U?P\d\s+)?.*?(0x[0-9a-f]+)", t)
        if m and "BRA" in t:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    for lo, hi in loops:
        body = ins[lo:hi + 1]
        hist = collections.Counter(opcode(t) for _, t in body)
        hm = hist.get("HMMA", 0)
        if hm == 0:
            continue
        tiles = hm / 4
        print(f"\nloop 0x{ins[lo][0]:x}..0x{ins[hi][0]:x}: {len(body)} instr, {hm} HMMA = {tiles:g} super-tiles -> "
              f"{len(body) / tiles:.1f} instr / super-tile, {len(body) / tiles / 16:.2f} / weight pair")
        alu = sum(v for k, v in hist.items() if k in ("SHF", "LOP3", "IADD3", "PRMT", "LEA", "ISETP", "SEL", "MOV", "VIADD", "IADD", "SGXT", "BMSK", "VIMNMX"))
        fma = sum(v for k, v in hist.items() if k in ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HMUL2", "HADD2"))
        print(f"  alu-pipe {alu} ({alu / tiles:.1f}/st)  fma-pipe {fma} ({fma / tiles:.1f}/st)  "
              + "  ".join(f"{k} {v / tiles:.1f}" for k, v in hist.most_common()))
    if "--dump" in sys.argv:
        for a, t in ins:
            print(f"{a:06x}  {t}")


if __name__ == "__main__":
    main()
