#!/usr/bin/env python
"""Per-source-line stall samples of one kernel from an ncu report (no GPU needed).

    python tools/ncu_lines.py <report.ncu-rep> <kernel substring> [lib.so] [top N]

ncu's CSV source page is SASS-only; the line table comes from `nvdisasm --print-line-info` of the same
cubin (instruction i of the report = instruction i of the disassembly)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(lib, kernel):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
        cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-c", "--print-line-info", cubin], capture_output=True, text=True).stdout
    out, cur, on = [], (None, 0), False
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", ln):
            out.append(cur)
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "multimodal-baby_b200", "lib", "libcvcl_b200.so")
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    names = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    # pick the section whose kernel name matches
    sec = None
    for ni in names:
        if kernel in rows[ni][1]:
            sec = ni
            break
    h = [i for i in hdr_i if i > sec][0]
    hdr = rows[h]
    end = min([i for i in names if i > sec] + [len(rows)])
    inst = [dict(zip(hdr, r)) for r in rows[h + 1:end] if r and r[0].startswith("0x")]
    lines = sass_lines(lib, kernel)
    if len(lines) != len(inst):
        print("warning: %d SASS instructions in the report, %d in the library (rebuilt since the capture?)" % (len(inst), len(lines)))
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg, src_cache = {}, {}
    total = 0
    for d, loc in zip(inst, lines):
        n = int(d.get("# Samples") or 0)
        total += n
        a = agg.setdefault(loc, {"n": 0, "inst": 0})
        a["n"] += n
        a["inst"] += int(d.get("Instructions Executed") or 0)
        for k in stalls:
            a[k] = a.get(k, 0) + int(d.get(k) or 0)

    def src(loc):
        f, l = loc
        if f is None:
            return ""
        if f not in src_cache:
            for base in (os.path.join(ROOT, "multimodal-baby_b200", "csrc"), ROOT):
                pth = os.path.join(base, f)
                if os.path.exists(pth):
                    src_cache[f] = open(pth).read().splitlines()
                    break
            else:
                src_cache[f] = []
        s = src_cache[f]
        return s[l - 1].strip()[:90] if 0 < l <= len(s) else ""
    print("total samples %d over %d instructions" % (total, len(inst)))
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
        st = sorted(((a.get(k, 0), k[6:]) for k in stalls), reverse=True)[:3]
        print("%6d %5.1f%%  %s:%-4d %-90s %s" % (a["n"], 100.0 * a["n"] / max(total, 1), loc[0], loc[1], src(loc),
                                                " ".join("%s:%d" % (k, v) for v, k in st if v)))


if __name__ == "__main__":
    main()
