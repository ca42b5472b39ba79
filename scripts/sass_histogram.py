"""SASS instruction histogram of the library's kernels (no GPU needed): cuobjdump -sass grit_b200/libmsda_b200.so,
grouped by opcode, for kernels whose demangled name matches a pattern.

    python scripts/sass_histogram.py "msda_fwd_v5<float, (int)32, (int)4, (int)4, (int)4, (bool)0>" > profiles/r02_sass_fwd_v5.txt
"""
import collections
import re
import subprocess
import sys

LIB = "grit_b200/libmsda_b200.so"


def main():
    pats = sys.argv[1:]
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True,
                           text=True).stdout.splitlines()
    blocks = re.split(r"\n\s*Function : ", "\n" + txt)[1:]
    for mangled_block, name in zip(blocks, names):
        if not all(p in name for p in pats):
            continue
        ops = collections.Counter()
        total = 0
        for line in mangled_block.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
            if m:
                op = m.group(1)
                ops[op] += 1
                total += 1
        print(f"== {name}\n   {total} SASS instructions (static)")
        fam = collections.Counter()
        for op, n in ops.items():
            fam[op.split(".")[0]] += n
        print("   by family: " + ", ".join(f"{k} {v}" for k, v in fam.most_common()))
        print("   memory / sync opcodes:")
        for op, n in sorted(ops.items(), key=lambda kv: -kv[1]):
            if op.split(".")[0] in ("LDG", "STG", "RED", "REDG", "ATOM", "ATOMG", "ATOMS", "LDS", "STS", "SHFL", "BAR", "LDGSTS",
                                    "UTMALDG", "UBLKCP", "SYNCS", "REDUX", "LDSM", "MEMBAR", "ERRBAR", "CCTL", "LDC", "ULDC",
                                    "VOTE", "MATCH", "UTCBAR", "UTCMMA", "BRA", "EXIT", "CALL", "RET", "BSSY", "BSYNC", "WARPSYNC"):
                print(f"      {op:32s} {n}")
        print()


if __name__ == "__main__":
    main()
