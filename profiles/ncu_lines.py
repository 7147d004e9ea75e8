#!/usr/bin/env python
"""profiles/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTR [LIB.so] -- per-source-line summary of an ncu capture.

ncu's CSV source page is per SASS instruction without line numbers; this joins it (by instruction order) with
`nvdisasm -g` of the kernel taken from the in-tree library, and prints instructions executed / stall samples
aggregated per (file, line).  Read-only tooling for the notes under profiles/.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
    out = []
    for f in sorted(os.listdir(d)):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        cur, inside = None, False
        for ln in txt.splitlines():
            if ln.startswith("//---") and ".text." in ln:
                inside = kernel in ln
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.append((int(m.group(1), 16), cur, m.group(2).strip()))
        if out:
            break
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "ploidyfrost_b200", "libpfgpu.so")
    top = int(os.environ.get("TOP", "40"))
    csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(csvtxt)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = [dict(zip(hdr, r)) for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
    sl = sass_lines(lib, kernel)
    if len(sl) != len(body):
        sys.stderr.write(f"warning: {len(body)} profiled instructions vs {len(sl)} disassembled (library rebuilt since the capture?)\n")
    base = int(body[0]["Address"], 16)
    by_off = {o: (loc, txt) for o, loc, txt in sl}
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_i = tot_s = 0
    for r in body:
        off = int(r["Address"], 16) - base
        loc = by_off.get(off, (None, ""))[0] or ("?", 0)
        ie = int(r["Instructions Executed"] or 0)
        te = int(r["Thread Instructions Executed"] or 0)
        sm = int(r["# Samples"] or 0)
        a = agg[loc]
        a[0] += ie
        a[1] += te
        a[2] += sm
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                a[3][c[6:]] += v
        tot_i += ie
        tot_s += sm
    print(f"{kernel}: {tot_i} warp-instructions executed, {tot_s} stall samples, {len(body)} SASS instructions")
    print(f"{'inst%':>6} {'smpl%':>6} {'thr/inst':>8}  location                     top stalls")
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        st = ", ".join(f"{k}:{v * 100 // max(a[2], 1)}%" for k, v in a[3].most_common(3))
        print(f"{a[0] * 100 / max(tot_i, 1):6.2f} {a[2] * 100 / max(tot_s, 1):6.2f} {a[1] / max(a[0], 1):8.1f}  {loc[0]}:{loc[1]:<6} {st}")


if __name__ == "__main__":
    main()
