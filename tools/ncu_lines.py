#!/usr/bin/env python
"""Attribute the per-instruction counters of an ncu --set full --import-source on capture to CUDA source lines.

    python tools/ncu_lines.py capture.ncu-rep kernel_substring [top_n] [--phases file.cuh]

--phases: additionally bucket the lines of file.cuh by the last preceding "// ---- " marker comment (the kernels' phases) and print,
per bucket, instructions, stall samples and the three largest stall reasons.

ncu's CSV source page is SASS-only; the line of every SASS instruction comes from nvdisasm -g on the cubin extracted from the
in-tree libbmf_b200.so (same build as the capture), matched by instruction offset inside the kernel.  Prints, per source line
(innermost inlined location), warp instructions executed, stall samples and their share of the kernel."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(kernel_sub):
    so = os.path.join(ROOT, "binarymeshfitting_b200", "libbmf_b200.so")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    out, inside, cur = {}, False, None
    for ln in txt.splitlines():
        if ln.startswith("\t.section\t.text."):
            inside = kernel_sub in ln
            cur = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    phases_file = sys.argv[sys.argv.index("--phases") + 1] if "--phases" in sys.argv else None
    if phases_file in argv:
        argv.remove(phases_file)
    rep, ksub = argv[0], argv[1]
    top = int(argv[2]) if len(argv) > 2 else 40
    lines = sass_lines(ksub)
    r = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    hdr_i = next(i for i, x in enumerate(rows) if x and x[0] == "Address")
    hdr = rows[hdr_i]
    ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Thread Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    per = {}
    per_stall = {}
    tot_i = tot_s = 0
    for x in rows[hdr_i + 1:]:
        if len(x) <= ci or not x[0].startswith("0x"):
            continue
        a = int(x[0], 16)
        if base is None:
            base = a
        loc, _ = lines.get(a - base, (None, ""))
        n, s, t = int(x[ci] or 0), int(x[cs] or 0), int(x[ct] or 0)
        e = per.setdefault(loc, [0, 0, 0])
        e[0] += n; e[1] += s; e[2] += t
        st = per_stall.setdefault(loc, {})
        for ci2, h in stall_cols:
            v = int(x[ci2] or 0) if ci2 < len(x) else 0
            if v:
                st[h] = st.get(h, 0) + v
        tot_i += n; tot_s += s
    print("kernel %s: %d warp instructions, %d stall samples" % (ksub, tot_i, tot_s))
    print("%-22s %12s %6s %9s %6s %5s  source" % ("file:line", "warp inst", "%", "samples", "%", "thr"))
    src_cache = {}
    for loc, (n, s, t) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
        text = ""
        if loc:
            f = os.path.join(ROOT, "binarymeshfitting_b200", "csrc", loc[0])
            if f not in src_cache:
                src_cache[f] = open(f).read().splitlines() if os.path.exists(f) else []
            if 0 < loc[1] <= len(src_cache[f]):
                text = src_cache[f][loc[1] - 1].strip()[:90]
        st = sorted(per_stall.get(loc, {}).items(), key=lambda kv: -kv[1])[:2]
        why = ",".join("%s %d%%" % (h.replace("stall_", ""), round(100.0 * v / max(s, 1))) for h, v in st)
        print("%-22s %12d %6.2f %9d %6.2f %5.1f  [%s] %s" % ("%s:%d" % loc if loc else "?", n, 100.0 * n / max(tot_i, 1), s, 100.0 * s / max(tot_s, 1), t / max(n, 1), why, text[:70]))


    if phases_file:
        src = open(os.path.join(ROOT, "binarymeshfitting_b200", "csrc", phases_file)).read().splitlines()
        marks = [(i + 1, l.strip()[:70]) for i, l in enumerate(src) if l.strip().startswith("// ---- ")]
        buckets = {}
        for loc, (n, s_, t) in per.items():
            key = "(other files / helpers)"
            if loc and loc[0] == phases_file:
                prev = [m for m in marks if m[0] <= loc[1]]
                key = prev[-1][1] if prev else "(before the first marker)"
            b = buckets.setdefault(key, [0, 0, {}])
            b[0] += n; b[1] += s_
            for h, v in per_stall.get(loc, {}).items():
                b[2][h] = b[2].get(h, 0) + v
        print("\nby phase of %s:" % phases_file)
        for key, (n, s_, st) in sorted(buckets.items(), key=lambda kv: -kv[1][1]):
            top3 = ", ".join("%s %.0f%%" % (h.replace("stall_", ""), 100.0 * v / max(s_, 1)) for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print("  %5.1f%% inst %5.1f%% samples  [%s]  %s" % (100.0 * n / max(tot_i, 1), 100.0 * s_ / max(tot_s, 1), top3, key))


if __name__ == "__main__":
    main()
