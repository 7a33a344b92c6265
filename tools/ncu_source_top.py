#!/usr/bin/env python3
"""Condense an `ncu --page source --csv --print-source sass` export: instruction mix, total stall samples, and the
instructions with the most samples.  Usage: python tools/ncu_source_top.py src.csv"""
import collections
import csv
import sys

csv.field_size_limit(1 << 30)
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
for i, r in enumerate(rows):
    if any("Source" == c or c == "Address" for c in r):
        hdr, rows = r, rows[i + 1:]
        break
if hdr is None:
    print("no header found; first rows:", rows[:3])
    sys.exit(0)
print("columns:", hdr)
col = {c: i for i, c in enumerate(hdr)}
src = col.get("Source")
samp = next((col[c] for c in hdr if c.startswith("# Samples") or c.startswith("Warp Stall Sampling (All")), None)
execd = next((col[c] for c in hdr if c.startswith("# Instructions Executed") or c.startswith("Instructions Executed")), None)
tot = 0
mix = collections.Counter()
per = []
for r in rows:
    if len(r) <= max(x for x in (src, samp, execd) if x is not None):
        continue
    try:
        s = float(r[samp].replace(",", "")) if samp is not None and r[samp] else 0.0
        e = float(r[execd].replace(",", "")) if execd is not None and r[execd] else 0.0
    except ValueError:
        continue
    op = r[src].split()[0] if r[src].split() else "?"
    if op.startswith("@"):
        op = r[src].split()[1] if len(r[src].split()) > 1 else op
    mix[op.rstrip(";")] += e
    tot += s
    per.append((s, e, r[src][:100]))
print("total samples", tot, "instructions executed (warp-level)", sum(mix.values()))
print("executed mix:", ", ".join("%s %d" % kv for kv in mix.most_common(24)))
stall_by_op = collections.Counter()
for s, e, t in per:
    op = t.split()[0] if t.split() else "?"
    if op.startswith("@") and len(t.split()) > 1:
        op = t.split()[1]
    stall_by_op[op.rstrip(";")] += s
print("samples by opcode:", ", ".join("%s %d" % kv for kv in stall_by_op.most_common(16)))
per.sort(reverse=True)
for s, e, t in per[:40]:
    print("%8.0f samples  %10.0f exec  %s" % (s, e, t))
