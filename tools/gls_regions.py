"""Where the GLS multifrontal kernel spends its time: aggregates the per-line warp-instruction counts and stall
samples of an `ncu --set full --import-source on` capture over the phases of mf_node (k2_gls.cu).
usage: python tools/gls_regions.py REPORT.ncu-rep"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
text = open("ninpol_b200/csrc/k2_gls.cu").read().splitlines()


def line_of(marker, start=0):
    for i in range(start, len(text)):
        if marker in text[i]:
            return i + 1
    raise KeyError(marker)


marks = [("setup: element + face rows", "// ---- setup: esup row, element groups ----"),
         ("adjacency bitmasks", "// ---- adjacency bitmasks"),
         ("leaf fronts (lane per front)", "// ---- leaf fronts, one per LANE ----"),
         ("flush of surviving original rows", "// ---- original groups that survived the leaf phase"),
         ("pivot choice + S list + chain decision", "// ---- elimination ----"),
         ("front assembly (cp.async)", "// (d) assemble one chunk"),
         ("panel factorisation", "// (e) panel: Householder on the three pivot columns"),
         ("reflector passes", "// (f) apply the three reflections to the other columns"),
         ("contribution block / R rows out", "// (g) rows below the pivot rows"),
         ("back substitution", "// ---- back substitution through the R rows"),
         ("weights + reroute test", "// ---- residual on the element rows, weights, CSR values ----"),
         ("end", "// persistent: one warp per CTA")]
bounds = [(name, line_of(m)) for name, m in marks]
cur, agg = None, {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
        continue
    try:
        ln, inst, samp = int(r[0]), int(r[7]), int(r[4])
    except ValueError:
        continue
    if cur == "k2_gls.cu":
        region = "helpers (hh_scalars, store_cb, compact, warp_or64 ...)"
        for (name, a), (_n, b) in zip(bounds[:-1], bounds[1:]):
            if a <= ln < b:
                region = name
                break
    else:
        region = "inlined intrinsics (%s)" % cur
    a = agg.setdefault(region, [0, 0])
    a[0] += inst
    a[1] += samp
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"total warp-instructions {ti}, stall samples {ts}")
for region, (inst, samp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{inst / ti * 100:5.1f}% inst {samp / ts * 100:5.1f}% samples  {region}")
