#!/usr/bin/env python
"""Per-source-line totals of an `ncu --page source --csv` SASS table: the i-th SASS row is matched with the i-th instruction of the
kernel in `nvdisasm -g -c` output of the library's cubin (line info from -lineinfo).
usage: sass_by_line.py <source.csv> <nvdisasm.txt> <mangled-kernel-substring> [top]"""
import csv, re, sys, collections
src_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
kern_plain = sys.argv[5] if len(sys.argv) > 5 else ""  # substring of the demangled name in the csv (several kernels in one report)
lines = open(dis).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
insn_line = []
cur = None
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        insn_line.append(cur)
rows = list(csv.reader(open(src_csv)))
# one section per profiled launch: "Kernel Name" line, header line, one row per SASS instruction; the first launch of the kernel is used
sec = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kern_plain in r[1]]
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address" and i > sec[0]][0]
h = rows[hdr]
ci, cs = h.index("Instructions Executed"), h.index("# Samples")
body = []
for r in rows[hdr + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
assert len(body) == len(insn_line), (len(body), len(insn_line))
inst, samp = collections.Counter(), collections.Counter()
for r, ln in zip(body, insn_line):
    inst[ln] += int(r[ci]); samp[ln] += int(r[cs])
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp instructions %d, samples %d" % (ti, ts))
for ln, v in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
    print("%-22s %6s  inst %5.1f%%  samples %5.1f%%" % (ln[0] if ln else "?", ln[1] if ln else "", 100.0 * v / ti, 100.0 * samp[ln] / max(1, ts)))
