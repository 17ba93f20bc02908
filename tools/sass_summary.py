"""Opcode histogram per kernel of the in-tree libsr4d.so (cuobjdump -sass; runs without a GPU).
    python tools/sass_summary.py [> profiles/rNN_sass_summary.txt]
Lists, per kernel, the instruction count and the counts of the mnemonics that prove the Blackwell path
(B200_PROFILING.md): UTCHMMA (tcgen05.mma kind::f16), UTCBAR (tcgen05.commit), LDTM / STTM (tcgen05.ld / st),
UTMALDG (cp.async.bulk.tensor = TMA tile loads), UBLKCP (cp.async.bulk), SYNCS (mbarrier), plus HMMA / FFMA / RED
to tell tensor-core kernels from CUDA-core ones."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "4dflownet_b200", "libsr4d.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "RED", "ATOMG", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[cur][k] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    arch = re.findall(r"arch = (sm_\w+)", out)
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, arch {sorted(set(arch))}; cuobjdump -sass opcode counts")
    print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{k:>7s}" for k in KEYS))
    tot = collections.Counter()
    for (name, c), dn in zip(kernels.items(), demangled):
        short = dn.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][-70:]
        print(f"{short:70s} {c['_total']:7d} " + " ".join(f"{c[k]:7d}" for k in KEYS))
        tot.update(c)
    print(f"{'TOTAL':70s} {tot['_total']:7d} " + " ".join(f"{tot[k]:7d}" for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())
