"""TEST / MEASUREMENT INFRASTRUCTURE without any checker in it: where the compiled unmodified reference lives (oracle/_ref, built by
`make -C oracle ref_full`) and how the reference's `-t N` output files are compared as schedule-independent multisets.  Imported by
tests/e2e_rows.py (which adds the oracle-based row regeneration) and by integration/time_program.py (which must not load the oracle)."""
from __future__ import annotations

import os

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_binaries():
    """(PloidyFrost, Bifrost) of oracle/_ref (built by `make -C oracle ref_full` in the dev container; they travel to the GPU box
    with the snapshot and need nothing from /root/reference at run time), or None."""
    ref = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
    pf, bf = os.path.join(ref, "PloidyFrost"), os.path.join(ref, "Bifrost")
    if not (os.path.exists(pf) and os.path.exists(bf)) and os.path.isdir("/root/reference/src"):   # dev container: build them
        import subprocess
        subprocess.run(["make", "-C", os.path.join(os.path.dirname(HERE), "oracle"), "ref_full"], check=False, capture_output=True)
    return (pf, bf) if os.path.exists(pf) and os.path.exists(bf) else None


def unitig_seq(d):
    return {ln.split("\t")[0]: ln.rstrip("\n").split("\t")[1] for ln in open(os.path.join(d, "P_Unitig_Id.txt"))}


def thread_dialect_view(d):
    """The `-t N` files as the schedule-independent things they are (row order, VarIds and unitig ids depend on
    the thread schedule): coverage rows without the VarId column, frequency lines, aligned bubbles keyed by the entrance / exit
    unitig SEQUENCES -- all as sorted multisets."""
    useq = unitig_seq(d)
    view = {}
    for a in ("bi", "tri", "tetra", "penta"):
        rows = []
        for ln in open(os.path.join(d, f"P_{a}cov.txt")):
            p = ln.rstrip("\n").split("\t")
            del p[-4]                                    # ... type, indelLen, VarId, VarNum, VarDis, ''
            rows.append("\t".join(p))
        view[a + "cov"] = sorted(rows)
        view[a + "fre"] = sorted(open(os.path.join(d, f"P_{a}fre.txt")).read().split("\n"))
    view["allfre"] = sorted(open(os.path.join(d, "P_allele_frequency.txt")).read().split("\n"))
    groups, ids = {}, []
    for ln in open(os.path.join(d, "P_alignseq.txt")):
        p = ln.rstrip("\n").split("\t")
        if p[0] not in groups:
            ids.append(int(p[0]))
        groups.setdefault(p[0], []).append((p[1], useq[p[2]], useq[p[3]], p[4]))
    view["alignseq"] = sorted(tuple(g) for g in groups.values())
    return view, ids


def colored_thread_dialect_view(d):
    """The coloured `-t N` files as sorted multisets (rows without the schedule-dependent VarId: ..., colour, type, indelLen, VarId,
    VarNum, CramerV, VarDis, '')."""
    useq = unitig_seq(d)
    view = {}
    for a in ("bi", "tri", "tetra", "penta"):
        rows = []
        for ln in open(os.path.join(d, f"P_{a}cov.txt")):
            p = ln.rstrip("\n").split("\t")
            del p[-5]
            rows.append("\t".join(p))
        view[a + "cov"] = sorted(rows)
        view[a + "fre"] = sorted(open(os.path.join(d, f"P_{a}fre.txt")).read().split("\n"))
    view["allfre"] = sorted(open(os.path.join(d, "P_allele_frequency.txt")).read().split("\n"))
    groups, ids = {}, []
    for ln in open(os.path.join(d, "P_alignseq.txt")):
        p = ln.rstrip("\n").split("\t")
        if p[0] not in groups:
            ids.append(int(p[0]))
        groups.setdefault(p[0], []).append((p[1], useq[p[2]], useq[p[3]], p[4]))
    view["alignseq"] = sorted(tuple(g) for g in groups.values())
    return view, ids

