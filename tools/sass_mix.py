#!/usr/bin/env python
"""Instruction mix of the innermost loop(s) of one kernel in a built library.

    python tools/sass_mix.py <lib.so> '<demangled-name substring>' [--dump]

Finds backward branches in the kernel's SASS, takes each loop body [target, branch] and prints
an opcode histogram grouped by pipe, plus instructions per fine step (loop handles 4 steps).
Static counts only; the dynamic numbers come from ncu (profiles/).
"""
import collections
import re
import subprocess
import sys

PIPES = [
    ("fp64", r"^(DFMA|DADD|DMUL|DSETP|DMNMX)"),
    ("fma/imad", r"^(FFMA|FMUL|FADD|IMAD|HFMA2)"),
    ("alu", r"^(LOP3|IADD3|IADD|SHF|PRMT|ISETP|FSETP|SEL|FSEL|MOV|LEA|FMNMX|LOP|PLOP3|VIADD|IABS|FLO|POPC|I2FP|VIMNMX|UMOV|FCHK|R2P|P2R|FSWZADD|IMNMX)"),
    ("xu(mufu/cvt)", r"^(MUFU|I2F|F2F|F2I|FRND)"),
    ("uniform", r"^(U[A-Z0-9]+|R2UR|S2UR|LDCU)"),
    ("mem", r"^(LD|ST|ATOM|RED|LDC|LDS|STS|LDG|STG|LDL|STL)"),
    ("ctrl", r"^(BRA|BSSY|BSYNC|EXIT|WARPSYNC|BAR|CALL|RET|NOP|YIELD|BREAK|BMOV)"),
]


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, body, res = None, [], {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                res[cur] = body
            cur, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if cur:
        res[cur] = body
    return res


def hot_loop_mix(lib, pat, steps_per_iter=4):
    """-> dict for the hottest loop of the first kernel whose demangled name contains `pat`:
    instruction count per 4 fine steps, split by dispatch class, and the 2-cycle-weighted figure
    (ALU-class and IMAD occupy the dispatch port for 2 cycles on B200, everything else 1)."""
    ks = kernels(lib)
    names = subprocess.run(["c++filt"] + list(ks), capture_output=True, text=True).stdout.splitlines()
    for mangled, name in zip(ks, names):
        if pat not in name:
            continue
        body = ks[mangled]
        # candidate loops = [target, backward branch]; the hot loop is the INNERMOST one (no other
        # backward branch inside, no global stores) that carries the most Philox multiplies (IMAD.WIDE)
        spans = []
        for addr, ins in body:
            m = re.search(r"BRA(?:\.U)?\s+(?:\S+,\s+)?(0x[0-9a-f]+)", ins)
            if m and int(m.group(1), 16) <= addr:
                spans.append((int(m.group(1), 16), addr))
        best, best_key = None, None
        for lo, hi in spans:
            if any((l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi for l2, h2 in spans):
                continue
            ops = [re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for a, i in body if lo <= a <= hi]
            if any(o.startswith("STG") for o in ops):
                continue              # the coarsest-level loop writes optional per-path outputs; the step loop stores nothing
            key = (sum(1 for o in ops if o.startswith("IMAD.WIDE")), len(ops))
            if key[0] and (best_key is None or key > best_key):
                best, best_key = ops, key
        if best is None:
            return None
        cls = collections.Counter()
        for o in best:
            for pipe, rx in PIPES:
                if re.match(rx, o):
                    cls[pipe] += 1
                    break
            else:
                cls["other"] += 1
        two = cls["alu"] + sum(1 for o in best if o.startswith("IMAD"))
        k = 4.0 / steps_per_iter
        return {"kernel": name.split("(")[0], "steps_per_loop_iteration": steps_per_iter,
                "inst_per_4_steps": len(best) * k, "classes_per_iteration": dict(cls),
                "dispatch_cycles_per_4_steps": (len(best) + two) * k}
    return None


def main():
    if "--json" in sys.argv:
        import json
        print(json.dumps(hot_loop_mix(sys.argv[1], sys.argv[2])))
        return
    lib, pat = sys.argv[1], sys.argv[2]
    dump = "--dump" in sys.argv
    ks = kernels(lib)
    names = subprocess.run(["c++filt"] + list(ks), capture_output=True, text=True).stdout.splitlines()
    for mangled, name in zip(ks, names):
        if pat not in name:
            continue
        body = ks[mangled]
        print("==", name.split("(")[0], f"({len(body)} instrs)")
        loops = []
        for addr, ins in body:
            m = re.search(r"BRA(?:\.U)?\s+(?:\S+,\s+)?(0x[0-9a-f]+)", ins)
            if m and int(m.group(1), 16) <= addr:
                loops.append((int(m.group(1), 16), addr))
        for lo, hi in loops:
            ins = [i for a, i in body if lo <= a <= hi]
            ops = [re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for i in ins]
            if not any(o.startswith(("DFMA", "FFMA")) for o in ops):
                continue
            hist = collections.Counter()
            full = collections.Counter(o.split(".")[0] for o in ops)
            for o in ops:
                for pipe, rx in PIPES:
                    if re.match(rx, o):
                        hist[pipe] += 1
                        break
                else:
                    hist["other:" + o.split(".")[0]] += 1
            print(f"  loop 0x{lo:x}-0x{hi:x}: {len(ins)} instrs = {len(ins) / 4:.1f}/step  ", dict(hist))
            print("     ", dict(full.most_common()))
            if dump:
                for a, i in body:
                    if lo <= a <= hi:
                        print(f"      {a:05x}  {i}")


if __name__ == "__main__":
    main()
