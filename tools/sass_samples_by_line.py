"""Join the per-instruction stall samples of an ncu report (`ncu -i X.ncu-rep --page source --csv`) with the line table of the kernel
(`nvdisasm -gi` of the cubin extracted with `cuobjdump -xelf all <object>`) and print samples per source line / line range.
Usage: python tools/sass_samples_by_line.py <samples.csv> <nvdisasm.txt> <mangled kernel name> [lo-hi:label ...]"""
import csv
import re
import sys
from collections import defaultdict

BODY_LINE = 328  # first line of the kernel body in qp_tile.cu (helpers above it are attributed to their call sites)


def main():
    csv_path, dis_path, kern = sys.argv[1:4]
    ranges = []
    for a in sys.argv[4:]:
        r, label = a.split(":")
        lo, hi = r.split("-")
        ranges.append((int(lo), int(hi), label))
    lines = open(dis_path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
    off2line = {}
    cur, chain, fresh = None, [], True
    for l in lines[start + 1:]:
        if l.startswith("//-----") or l.startswith("\t.section"):
            break
        if "//## File" in l:
            # `nvdisasm -gi` prints the inlining chain innermost first; attribute the instruction to the first frame that lies in the
            # kernel body (or in a lambda of it), i.e. in the .cu file at or behind BODY_LINE
            if fresh:
                chain, fresh = [], False
            chain += [(f.split("/")[-1], int(n)) for f, n in re.findall(r'"([^"]+)", line (\d+)', l)]
            body = [c for c in chain if c[0].endswith(".cu") and c[1] >= BODY_LINE]
            cur = body[0] if body else chain[0]
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", l)
        if m and cur:
            off2line[int(m.group(1), 16)] = (cur, m.group(2))
            fresh = True
    rows = list(csv.reader(open(csv_path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    si, ei = hdr.index("# Samples"), hdr.index("Instructions Executed")
    body = rows[hdr_i + 1:]
    base = int(body[0][0], 16)
    per_line = defaultdict(lambda: [0, 0])
    total = 0
    for r in body:
        off = int(r[0], 16) - base
        (f, ln), _ = off2line.get(off, (("?", 0), ""))
        s, e = int(r[si]), int(r[ei])
        per_line[(f, ln)][0] += s
        per_line[(f, ln)][1] += e
        total += s
    print("total samples", total)
    if ranges:
        acc = defaultdict(lambda: [0, 0])
        for (f, ln), (s, e) in per_line.items():
            lab = "other"
            if f == "qp_tile.cu" or f.endswith(".cu"):
                for lo, hi, label in ranges:
                    if lo <= ln <= hi:
                        lab = label
                        break
            else:
                lab = "other:" + f
            acc[lab][0] += s
            acc[lab][1] += e
        for lab, (s, e) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
            print("%-40s %8d samples %5.1f %%   %12d warp-instructions" % (lab, s, 100.0 * s / total, e))
    else:
        for (f, ln), (s, e) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:60]:
            print("%-24s %5d %8d %5.1f %% %12d" % (f, ln, s, 100.0 * s / total, e))


if __name__ == "__main__":
    main()
