"""Per-pipe SASS instruction mix of a kernel's hottest loop (default: seed_scan_kernel's 16-positions-per-word loop).

    python tools/sass_mix.py [object] [kernel substring] [units per loop iteration]

Reads `cuobjdump -sass` of the object the library is linked from, cuts out the kernel, finds the innermost loop with
the most 64-bit multiplies (IMAD.WIDE) - for seed_scan_kernel that is one 16-base word = 16 positions, two exact
64-bit hashes each - and counts its instructions by the pipe that issues them on sm_100:
    alu   LOP3 SHF IADD3 ISETP SEL PRMT VIMNMX MOV LEA VIADD VABSDIFF ...   (INT32 / logic pipe, 16 lanes per sub-partition
                                                                           => one warp instruction every 2 cycles)
    fma   IMAD* (also IMAD.MOV / IMAD.SHL / IMAD.IADD, which the compiler uses to offload the ALU pipe)
    xu    POPC BREV FLO MUFU (special-function unit)
    lsu   LDS STS LDG STG ATOM RED
    ctl   BRA BSSY BSYNC EXIT WARPSYNC NOP ...
    uni   U* (uniform datapath)
This is the reproducible form of the "ALU / FMA instructions per base" figures in DESIGN.md section 4.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ALU = ("LOP3", "SHF", "IADD3", "IADD", "ISETP", "SEL", "PRMT", "VIMNMX", "IMNMX", "MOV", "LEA", "VIADD", "VABSDIFF", "IABS", "SGXT", "BMSK",
       "P2R", "R2P", "PLOP3", "CS2R", "S2R", "ICMP", "FSEL", "LOP")
XU = ("POPC", "BREV", "FLO", "MUFU")
LSU = ("LDS", "STS", "LDG", "STG", "LD", "ST", "ATOM", "ATOMS", "ATOMG", "RED", "LDC", "LDSM", "SHFL", "VOTE", "MATCH", "REDUX")
CTL = ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "NOP", "BAR", "RET", "CALL", "YIELD", "DEPBAR", "BREAK", "BMOV", "ERRBAR", "MEMBAR", "NANOSLEEP")


def pipe_of(op):
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UMOV_",):
        return "uni"
    if base.startswith("IMAD") or base in ("FFMA", "FMUL", "FADD", "IDP", "IDP4A"):
        return "fma"
    if base in XU:
        return "xu"
    if base in LSU:
        return "lsu"
    if base in CTL:
        return "ctl"
    if base in ALU:
        return "alu"
    return "other:" + base


def kernel_sass(obj, name):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    cur, keep, chosen = None, [], None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            if chosen is None and name in cur:
                chosen = cur               # the first match only: template instances share a name prefix
            continue
        if cur and cur == chosen:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                keep.append((int(m.group(1), 16), m.group(2).strip()))
    if not keep:
        raise SystemExit("no kernel matching %r in %s" % (name, obj))
    return keep


def main():
    obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pyskani_b200", "csrc", "seed_kernels.o")
    name = sys.argv[2] if len(sys.argv) > 2 else "seed_scan_kernelILb0"      # <false>: the high-word comparison variant
    units = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    ins = kernel_sass(obj, name)
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, text) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    if not loops:
        raise SystemExit("no loop found")

    def n_wide(lo, hi):
        return sum(1 for _, t in ins[lo:hi + 1] if "IMAD.WIDE" in t)
    # innermost = no other loop strictly inside; among those the one with the most IMAD.WIDE
    inner = [l for l in loops if not any(o != l and o[0] >= l[0] and o[1] <= l[1] for o in loops)]
    lo, hi = max(inner or loops, key=lambda l: n_wide(*l))
    body = ins[lo:hi + 1]
    counts, ops = collections.Counter(), collections.Counter()
    for _, text in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", text)
        op = t.split()[0]
        p = pipe_of(op)
        counts[p] += 1
        ops[(p, op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("IMAD.") and len(op.split(".")) > 1 else ""))] += 1
    total = sum(counts.values())
    print("# %s, loop at 0x%04x..0x%04x: %d instructions per iteration = %d units (positions) per iteration" % (name, ins[lo][0], ins[hi][0], total, units))
    print("# pipe   per iteration   per unit")
    for p, c in sorted(counts.items(), key=lambda kv: -kv[1]):
        print("%-8s %8d %12.2f" % (p, c, c / units))
    print("%-8s %8d %12.2f" % ("total", total, total / units))
    print("# opcodes")
    for (p, op), c in sorted(ops.items(), key=lambda kv: (-kv[1], kv[0])):
        print("%-6s %-14s %6d %8.2f" % (p, op, c, c / units))
    alu, fma = counts["alu"], counts["fma"]
    # issue model of one SM sub-partition: ALU pipe one warp instruction per 2 cycles; FMA pipe likewise for integer multiplies,
    # with IMAD.WIDE and IMAD.HI taking two slots each (measured: tools/micro/int_pipes.cu, profiles/r2_int_pipes.txt), and an
    # IMAD.WIDE costing the ALU pipe about one slot as well; the dispatch port issues one instruction per cycle
    wide = sum(c for (p, op), c in ops.items() if op == "IMAD.WIDE")
    high = sum(c for (p, op), c in ops.items() if op == "IMAD.HI")
    cyc = max(2 * (alu + wide), 2 * (fma + wide + high), total)
    print("# issue model per iteration: ALU %d cycles (+ %d for IMAD.WIDE), FMA %d cycles (IMAD.WIDE and IMAD.HI double), dispatch %d cycles"
          " -> bound %d cycles = %.2f cycles per unit per warp" % (2 * alu, 2 * wide, 2 * (fma + wide + high), total, cyc, cyc / units))
    print("# => %.1f G units/s per GPU at 148 SMs x 4 sub-partitions x 32 lanes x 1.965 GHz" % (148 * 4 * 32 * 1.965 / (cyc / units)))


if __name__ == "__main__":
    main()
