#!/usr/bin/env python
"""Warp-stall samples of one kernel of an .ncu-rep, aggregated per CUDA source line.

ncu's CSV source page carries the samples per SASS instruction only; the line table comes from
nvdisasm -g on the cubin extracted from libb2s.so (instruction order is identical).
usage: ncu_stalls.py <report.ncu-rep> <cubin-name e.g. conv_tc3> <mangled-name substring> [kernel-id]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, cubin_key, fn_key = sys.argv[1:4]
kid = sys.argv[4] if len(sys.argv) > 4 else "1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "openpcseg_b200", "libb2s.so")], cwd=td,
                   capture_output=True)
    cub = [f for f in os.listdir(td) if f.startswith(cubin_key + ".")][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
cur, infn, ins = None, False, []
for l in sass.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = fn_key in m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if infn:
        m = re.match(r"\s+(/\*[0-9a-f]+\*/)\s+(.*?);", l)
        if m:
            ins.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
idx = {n: i for i, n in enumerate(h)}
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
if len(body) >= 2 * len(ins):
    body = body[:len(body) // 2]
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg, st = collections.Counter(), collections.defaultdict(collections.Counter)
for j in range(min(len(ins), len(body))):
    r = body[j]
    agg[ins[j]] += int(r[idx["# Samples"]] or 0)
    for s in stalls:
        if r[idx[s]] and int(r[idx[s]]):
            st[ins[j]][s] += int(r[idx[s]])
tot = sum(agg.values())
print(f"# {len(ins)} SASS instructions, {len(body)} profiled rows, {tot} samples")
src = {}
for k, v in agg.most_common(28):
    f = os.path.join(root, "openpcseg_b200", "csrc", k[0]) if k else None
    text = ""
    if f and os.path.exists(f):
        src.setdefault(f, open(f).read().split("\n"))
        text = src[f][k[1] - 1].strip()[:70]
    print(f"{v:6d} {100.0 * v / tot:5.1f}%  {k[0]}:{k[1]:<4d} {dict(st[k].most_common(2))}  | {text}")
