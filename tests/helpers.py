"""Shared helpers of the parity tests (test infrastructure)."""
import hashlib

import numpy as np


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rot_err(Ra, Rb) -> float:
    """|skew(Ra^T Rb)| - robust for fp32 matrices (arccos((tr-1)/2) has a ~2e-4 rad noise floor)."""
    D = np.asarray(Ra, np.float64).reshape(3, 3).T @ np.asarray(Rb, np.float64).reshape(3, 3)
    return float(np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]]) / 2)


def canonical(recs: np.ndarray) -> np.ndarray:
    """Canonical sort + adjacent unique of a structured match array - numpy restatement for multiset checks."""
    if len(recs) == 0:
        return recs
    order = np.lexsort((recs["x"], recs["y"], recs["class_idx"], recs["template_id"], -recs["similarity"].astype(np.float64)))
    s = recs[order]
    keep = np.ones(len(s), bool)
    keep[1:] = ~((s["x"][1:] == s["x"][:-1]) & (s["y"][1:] == s["y"][:-1]) & (s["similarity"][1:] == s["similarity"][:-1]) &
                 (s["class_idx"][1:] == s["class_idx"][:-1]))
    # adjacency is evaluated on the ORIGINAL sorted sequence, like std::unique keeps the first of each run
    out = [s[0]]
    for i in range(1, len(s)):
        a = out[-1]
        if a["x"] == s[i]["x"] and a["y"] == s[i]["y"] and a["similarity"] == s[i]["similarity"] and a["class_idx"] == s[i]["class_idx"]:
            continue
        out.append(s[i])
    return np.array(out, dtype=recs.dtype)


def tset_from_npz(z, synth):
    T = tuple(int(t) for t in z["T"])
    nc = int(z["class_of"].max()) + 1 if len(z["class_of"]) else 0
    return synth.TemplateSet(len(T), 2, T, ["obj%02d" % c for c in range(nc)], z["headers"], z["features"], z["class_of"], z["pose13"])
