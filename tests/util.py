"""Shared comparison helpers for the parity tests."""
import numpy as np

MSA_KEYS = ("status", "n_rows", "aln_len", "rows_off", "rows", "var_off", "var_col", "var_kind", "cls_off", "cls",
            "ilen_off", "ilen")


def assert_msa_equal(a: dict, b: dict, bubbles=None, what=""):
    """Bit-exact comparison of two pf_msa_batch_t dumps; on mismatch shows the first differing bubble."""
    from oracle.bindings import msa_bubble
    for key in MSA_KEYS:
        x, y = np.asarray(a[key]), np.asarray(b[key])
        if x.shape != y.shape or not np.array_equal(x, y):
            detail = ""
            n = min(len(a["n_rows"]), len(b["n_rows"]))
            for i in range(n):
                try:
                    p, q = msa_bubble(a, i), msa_bubble(b, i)
                except Exception:
                    break
                if p != q:
                    detail = f"\nfirst differing bubble {i}: input={bubbles[i] if bubbles else '?'}\n expected={p}\n got={q}"
                    break
            raise AssertionError(f"{what}: field '{key}' differs{detail}")
