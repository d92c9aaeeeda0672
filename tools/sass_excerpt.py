"""Per-kernel SASS evidence of libsqp_b200.so (cuobjdump -sass): counts of the mnemonics that prove the hardware paths, one example each.
Usage: python tools/sass_excerpt.py > profiles/rNN_sass_excerpt.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sqp_solver_b200", "libsqp_b200.so")
KEYS = ["UBLKCP", "SYNCS", "DMMA", "UCGABAR", "DFMA", "DADD", "DMUL", "FFMA", "SHFL", "BAR.SYNC", "ATOMG", "LDS", "STS", "LDG", "STG", "LDL", "STL",
        "MUFU", "MEMBAR", "FSEL"]
EXAMPLES = ["UBLKCP", "SYNCS", "DMMA", "UCGABAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    print("# SASS evidence: `cuobjdump -sass sqp_solver_b200/libsqp_b200.so` (python tools/sass_excerpt.py), per kernel: instruction counts of the")
    print("# mnemonics that prove the hardware paths (B200_PROFILING.md: UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, DMMA = fp64")
    print("# tensor-core mma.sync m8n8k4, UCGABAR_* = barrier.cluster, DFMA/DADD/DMUL = fp64 pipe, SHFL = warp shuffles, ATOMG = work-queue atomics)")
    print("# and one example line of each. tcgen05 (UTC*MMA / LDTM) is absent on purpose: it has no fp64 kind, and every iteration-path")
    print("# contraction is a mat-vec with one right-hand side per matrix. FSEL: the select-free reduce-scatter trees of the register-tiled")
    print("# kernel removed 4 FSEL per fp64 value and tree step from the ADMM loop (round 2; compare profiles/r02_sass_excerpt.txt).")
    print("# arch:", ", ".join(sorted(set(re.findall(r"arch = (\S+)", sass)))))
    blocks = re.split(r"\n\s*Function : ", sass)[1:]
    for name, blk in zip(names, blocks):
        cnt = collections.Counter()
        ex = {}
        for line in blk.split("\n"):
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + ".") or op.startswith(k + "_") or (k == "UCGABAR" and op.startswith("UCGABAR")):
                    cnt[k] += 1
                    if k in EXAMPLES and k not in ex:
                        ex[k] = re.sub(r"\s*/\*.*", "", line.split("*/", 1)[1]).strip(" ;")
        print("\n## " + name)
        print("   " + "  ".join("%s=%d" % (k, cnt[k]) for k in KEYS if cnt[k]))
        for k in EXAMPLES:
            if k in ex:
                print("   e.g. " + ex[k])


if __name__ == "__main__":
    main()
