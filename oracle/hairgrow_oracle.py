"""TEST INFRASTRUCTURE — CPU oracle for the HairGrow strand trace.  NOT part of the product.

Restatement of HairGrowing.trace / traceFromScalp / GenerateGuideStrandFromScalp / randomlyGenerateSegments
(/root/reference/HairGrow.py:59-299) with scalar numpy float32 arithmetic in the reference's operation order
(torch.dot of 3-vectors = (a0*b0 + a1*b1) + a2*b2 with separately rounded products; norms accumulate with FMAs).
Random jitter is injected (one [3] row per trace call, in call order) so both sides consume identical draws.

Parity pin: tests/golden/hairgrow_small.npz, produced by the unmodified reference with torch.rand_like patched to
read the same injected rows (tests/golden/make_golden_hairgrow.py).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _fma(a, b, c):
    return F(np.float64(a) * np.float64(b) + np.float64(c))


def dot3(a, b):
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def norm3(a):
    return F(np.sqrt(_fma(a[2], a[2], _fma(a[1], a[1], F(a[0] * a[0])))))


class Volume:
    """occ [Z,Y,X], ori [3,Z,Y,X] float32 in HairGrowing's frame (ori[1:] already negated, HairGrow.py:55)."""

    def __init__(self, occ, ori):
        self.occ = np.ascontiguousarray(occ, dtype=F)
        self.ori = np.ascontiguousarray(ori, dtype=F)
        self.Z, self.H, self.W = self.occ.shape

    @classmethod
    def from_memory(cls, occ_xyz, ori_xyz3):
        """from PMVO.refine's in-memory arrays occ [X,Y,Z], ori [X,Y,Z,3] (world signs)."""
        occ = np.transpose(occ_xyz, (2, 1, 0)).astype(F)
        ori = np.transpose(ori_xyz3, (3, 2, 1, 0)).astype(F).copy()
        ori[1:] *= -1
        return cls(occ, ori)

    def idx(self, p):
        # .type(torch.long): truncation toward zero, then clamp (HairGrow.py:66-69)
        x = min(max(int(p[0]), 0), self.W - 1)
        y = min(max(int(p[1]), 0), self.H - 1)
        z = min(max(int(p[2]), 0), self.Z - 1)
        return z, y, x

    def tan(self, i):
        return self.ori[:, i[0], i[1], i[2]].copy()


def _walk(vol, seed, sign, thr, max_steps=256):
    """one direction of trace (HairGrow.py:78-105 / :116-143)."""
    pos = seed.copy()
    i = vol.idx(pos)
    tan = vol.tan(i)
    out = []
    count = 0
    while True:
        if vol.occ[i] == 0:
            break
        nxt = (pos + sign * tan).astype(F)
        ni = vol.idx(nxt)
        ntan = vol.tan(ni)
        if dot3(ntan, tan) < F(thr):
            break
        pos, tan, i = nxt, ntan, ni
        out.append(pos.copy())
        count += 1
        if count >= max_steps:
            break
    return out


def trace(vol, seed_row, flag, thr, jitter):
    """HairGrowing.trace (HairGrow.py:59-149).  seed_row is modified IN PLACE (§9-R8).  Returns [L,3] or None."""
    seed_row += F(0.5)
    seed_row += (jitter.astype(F) * F(0.5)).astype(F)
    seed = seed_row.copy()
    if flag[vol.idx(seed)] >= 3:
        return None
    fwd = _walk(vol, seed, F(1.0), thr)
    bwd = _walk(vol, seed, F(-1.0), thr)
    strand = bwd[::-1] + [seed] + fwd
    if len(strand) >= 5:
        return np.stack(strand).astype(F)
    return None


def trace_from_scalp(vol, root, normal, thr, max_steps=256, max_inner=25):
    """HairGrowing.traceFromScalp (HairGrow.py:154-223)."""
    pos = root.astype(F).copy()
    n = normal.astype(F)
    d = np.array([0, 1, 0], F)
    m = min(F(dot3(n, d) + F(1.0)), F(1.0))
    tan = (n + d * m).astype(F)
    tan = (tan / norm3(tan)).astype(F)
    strand = [pos.copy()]
    count, inner = 0, True
    i = vol.idx(pos)
    while True:
        if vol.occ[i] == 0 and not inner:
            break
        nxt = (pos + tan).astype(F)
        ni = vol.idx(nxt)
        ntan = vol.tan(ni)
        if norm3(ntan) < F(0.1) and inner:
            if dot3(tan, n) < F(0.85):
                ntan = tan.copy()
            else:
                ntan = (tan + d * m).astype(F)
                ntan = (ntan / norm3(ntan)).astype(F)
        else:
            if dot3(ntan, tan) < F(thr) and not inner:
                if dot3(-ntan, tan) < F(thr):
                    break
                ntan = -ntan
            if dot3(ntan, tan) < 0 and inner:
                ntan = -ntan
            inner = False
        pos, tan, i = nxt, ntan, ni
        strand.append(pos.copy())
        count += 1
        if count >= max_steps:
            break
        if count >= max_inner and inner:
            break
    return None if inner else np.stack(strand).astype(F)


def _bump(vol, flag, strand, mode):
    idx = np.array([vol.idx(p) for p in strand])
    if mode == 1:
        flag[idx[:, 0], idx[:, 1], idx[:, 2]] = 1
    else:
        flag[idx[:, 0], idx[:, 1], idx[:, 2]] += 1          # numpy fancy += : once per unique voxel, like torch


def positive_seeds(vol):
    """torch.nonzero(occ) order (z,y,x) flipped to (x,y,z) float (HairGrow.py:230-232)."""
    nz = np.argwhere(vol.occ != 0)
    return nz[:, ::-1].astype(F).copy()


def generate_guide_strands(vol, scalp_points, scalp_normals, thr, jitter, passes=2):
    """GenerateGuideStrandFromScalp (HairGrow.py:226-265).  jitter [passes*M,3] in call order."""
    seeds = positive_seeds(vol)
    flag = np.zeros_like(vol.occ)
    strands = []
    for r, nrm in zip(scalp_points, scalp_normals):
        s = trace_from_scalp(vol, r, nrm, thr)
        if s is not None:
            strands.append(s)
            _bump(vol, flag, s, 1)
    num_root = len(strands)
    strands += segments_passes(vol, seeds, flag, thr, jitter, passes)
    return strands, num_root


def segments_passes(vol, seeds, flag, thr, jitter, passes):
    out, c = [], 0
    for _ in range(passes):
        for i in range(seeds.shape[0]):
            s = trace(vol, seeds[i], flag, thr, jitter[c])
            c += 1
            if s is not None:
                out.append(s)
                _bump(vol, flag, s, 0)
    return out


def randomly_generate_segments(vol, thr, jitter, passes=3):
    """randomlyGenerateSegments (HairGrow.py:269-299)."""
    return segments_passes(vol, positive_seeds(vol), np.zeros_like(vol.occ), thr, jitter, passes)


# ---------------------------------------------------------------------------------------------------------------------
# Strand smoothing (Utils/Utils.py:1148-1198).  TEST INFRASTRUCTURE like the rest of this file.
def smooth_strand(strand, lap_constraint=2.0, pos_constraint=1.0, fix_tips=False):
    """smnooth_strand (Utils.py:1148-1192): least squares of [lap*L ; pos*I] x = [0 ; pos*s] per axis through the normal
    equations, solved with scipy's sparse LU in float64 and stored back into an array of the strand's dtype.
    L: rows (1,-1), (-1,2,-1) x (n-2), (-1,1)."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import spsolve
    n = strand.shape[0]
    rows, cols, vals = [0, 0], [0, 1], [1.0, -1.0]
    for k in range(1, n - 1):
        rows += [k, k, k]; cols += [k - 1, k, k + 1]; vals += [-1.0, 2.0, -1.0]
    rows += [n - 1, n - 1]; cols += [n - 2, n - 1]; vals += [-1.0, 1.0]
    vals = [v * lap_constraint for v in vals]
    rows += list(range(n, 2 * n)); cols += list(range(n)); vals += [pos_constraint] * n
    A = sp.coo_matrix((np.array(vals), (np.array(rows), np.array(cols))), shape=(2 * n, n))
    At = A.transpose()
    AtA = At.dot(A)
    out = np.copy(strand)
    for a in range(3):
        b = np.zeros(2 * n)
        b[n:] = out[:, a] * pos_constraint                 # float32 strand: the product is rounded to float32 (Utils.py:1179)
        out[:, a] = spsolve(AtA, At.dot(b))[:n]
    if fix_tips:
        res = strand.copy()
        res[1:-1] = out[1:-1]
        return res
    return out


def smooth_strand_banded(strand, lap_constraint=2.0, pos_constraint=1.0):
    """The same system solved the way csrc/smooth.cu does (bands of A^T A by accumulated row outer products, banded
    Cholesky, float64): used on CPU to check that formulation against smooth_strand before it is trusted on the GPU."""
    s = np.asarray(strand)
    n = s.shape[0]
    l2, p2 = float(lap_constraint) ** 2, float(pos_constraint) ** 2
    D, E, F = np.full(n, p2), np.zeros(n), np.zeros(n)
    D[0] += l2; D[1] += l2; E[0] += -l2
    for k in range(1, n - 1):
        D[k - 1] += l2; D[k] += 4 * l2; D[k + 1] += l2
        E[k - 1] += -2 * l2; E[k] += -2 * l2; F[k - 1] += l2
    D[n - 2] += l2; D[n - 1] += l2; E[n - 2] += -l2
    for j in range(n):
        t = D[j]
        if j >= 1: t -= E[j - 1] ** 2
        if j >= 2: t -= F[j - 2] ** 2
        D[j] = np.sqrt(t)
        if j + 1 < n:
            u = E[j]
            if j >= 1: u -= F[j - 1] * E[j - 1]
            E[j] = u / D[j]
        if j + 2 < n: F[j] = F[j] / D[j]
    X = np.zeros((n, 3))
    rhs = float(pos_constraint) * (s.astype(np.float32) * np.float32(pos_constraint)).astype(np.float64)
    for j in range(n):
        y = rhs[j].copy()
        if j >= 1: y -= E[j - 1] * X[j - 1]
        if j >= 2: y -= F[j - 2] * X[j - 2]
        X[j] = y / D[j]
    for j in range(n - 1, -1, -1):
        x = X[j].copy()
        if j + 1 < n: x -= E[j] * X[j + 1]
        if j + 2 < n: x -= F[j] * X[j + 2]
        X[j] = x / D[j]
    return X.astype(s.dtype)
