"""Decision margins of the reference's discrete choices, computed with the CPU oracle (test infrastructure only).

The hot path ends in selected items -- a depth sample, a medoid, a voxel index -- so a flipped arg-min is an O(1)
difference, not a rounding error.  Parity is therefore stated as: EXACT agreement on every item whose decision margin
in the oracle exceeds eps, with the number of excluded items reported (SURVEY.md §7 "Hard parts").  The margins are
recorded next to the goldens (tests/golden/add_margins.py, make_golden_full.py) and recomputed nowhere on the product
path.
"""
import numpy as np
import torch

from . import pmvo_oracle as O


def forward_margins(vm, points, P, conf_threshold, eps=5e-8, with_threshold_gap=False):
    """PMVO.forward (PMVO.py:39-78): gap between the winning (base view, depth sample) loss and the runner-up over all
    (valid base, sample) pairs.  -> margin float64 [N] (inf when there is a single candidate).

    with_threshold_gap: also return, per point, how close a confidence test that feeds the choice was to flipping
    (PMVO.py:196-205): min |sum(w)/count - conf_threshold| over the (valid base, sample) pairs whose flip would matter --
    the winner itself (its loss turns into 1 / its high-confidence flag changes), a masked pair whose raw loss would
    win (raw < winning loss + eps), and every pair of a base whose positive count sits at the `< 5` boundary (4 or 5)."""
    _, _, _, _, dbg = O.forward(vm, points, P, conf_threshold, debug=True)
    L = torch.stack(dbg["L"], 0).double()                       # [10, N, S]
    valid = (dbg["base_conf"] > 0)                              # [10, N]; base 0 is taken unconditionally (PMVO.py:57-64)
    valid[0] = True
    L = torch.where(valid[:, :, None], L, torch.full_like(L, float("inf")))
    L = torch.where(torch.isnan(L), torch.full_like(L, -float("inf")), L)      # NaN wins torch.min: margin 0 below
    flat = L.permute(1, 0, 2).reshape(L.size(1), -1)
    two = torch.topk(flat, 2, dim=1, largest=False).values
    m = two[:, 1] - two[:, 0]
    m = torch.where(torch.isnan(m), torch.zeros_like(m), m)
    if not with_threshold_gap:
        return m.numpy()
    raw = torch.stack([d["raw"] for d in dbg["detail"]], 0).double()
    ratio = torch.stack([d["ratio"] for d in dbg["detail"]], 0).double()
    pos = torch.stack([d["pos"] for d in dbg["detail"]], 0)
    npos = pos.sum(-1)                                          # [10, N]
    win = two[:, 0][None, :, None]
    matters = (L <= win) | (~pos & (raw < win + eps)) | ((npos == 4) | (npos == 5))[:, :, None]
    matters &= valid[:, :, None]
    gap = torch.where(matters, (ratio - conf_threshold).abs(), torch.full_like(ratio, float("inf")))
    gap = torch.where(torch.isnan(gap), torch.zeros_like(gap), gap)
    return m.numpy(), gap.amin(dim=(0, 2)).numpy()


def singleton_base_groups(vm, points, P, chunk):
    """Points whose depth samples the reference computes on a batch of ONE (PMVO.py:290-296: sample_next_3d_pos projects
    `points[base_view == v]` per view).  When a point is the only one of its forward() chunk with view v at some used
    base rank, torch.matmul(pose, hom[4,1]) takes MKL's matrix-vector path, whose accumulation order differs from the
    sgemm path every other point sees (44 % of such projections move by an ulp) -- the reference's result for such a
    point depends on what else happens to be in its chunk, so it is reported separately, not as a kernel difference.
    -> bool [N]"""
    pts = torch.from_numpy(np.asarray(points)).type(torch.float)
    out = np.zeros(pts.size(0), dtype=bool)
    for lo in range(0, pts.size(0), chunk):
        sub = pts[lo:lo + chunk]
        st = O.compute_visible_and_ori(vm, sub, P)
        bidx, bval, _ = O.find_base_views(st["visible"], st["Conf"])
        for r, i in enumerate(range(0, 20, 2)):
            used = (bval[i] > 0) if r else torch.ones_like(bval[i], dtype=torch.bool)   # rank 0 is taken unconditionally
            cnt = torch.bincount(bidx[i], minlength=vm.V)
            out[lo:lo + chunk] |= ((cnt[bidx[i]] == 1) & used).numpy()
    return out


def _medoid_gap(ori_nk3):
    """compute_points_similarity (PMVO_utils.py:366-382): best mean |cos| minus second best.  [N,K,3] -> [N]"""
    N, K, _ = ori_nk3.size()
    a = ori_nk3[:, :, None, :].expand(-1, -1, K, -1)
    b = a.permute(0, 2, 1, 3)
    sim = torch.maximum(torch.cosine_similarity(a, b, dim=-1), torch.cosine_similarity(-a, b, dim=-1))
    mean = torch.mean(sim, dim=-1).double()
    if K < 2:
        return np.full((N,), np.inf)
    two = torch.topk(mean, 2, dim=1).values
    return (two[:, 0] - two[:, 1]).numpy()


def knn_gaps(tree, query, k):
    """gap between the k-th and (k+1)-th neighbour distance (set membership) and the smallest gap between consecutive
    neighbour distances (order: torch.mean sums in neighbour order).  -> (idx [n,k], gap [n])"""
    kk = min(k + 1, tree.n)
    d, nn = tree.query(query, kk)
    d = np.atleast_2d(d)
    gap = np.min(np.diff(d, axis=1), axis=1) if d.shape[1] > 1 else np.full((d.shape[0],), np.inf)
    return np.atleast_2d(nn)[:, :k], gap


def refine_margins(points, ori_in, ori_out_oracle, sub_num=5000, k=100):
    """PMVO.refine step (i) (PMVO.py:605-641), replayed on the oracle's own result: per point the kNN distance gap, the
    medoid gap of the neighbour orientations it saw (earlier chunks already updated: §9-R7) and the distance of
    |cos(center, ori)| from the 0.95 update threshold; `tainted` closes the ambiguity over the chunk order: a point
    that gathers an ambiguous (or tainted) point of an EARLIER chunk may legitimately differ too.
    -> dict(knn_gap, medoid_gap, update_gap [N], nbr [N,k])"""
    from scipy.spatial import KDTree
    points = np.asarray(points)
    n = points.shape[0]
    tree = KDTree(data=points)
    cur = np.array(ori_in, copy=True)
    kg, mg, ug = np.zeros(n), np.zeros(n), np.zeros(n)
    nbr = np.zeros((n, min(k, n)), dtype=np.int64)
    for i in range(n // sub_num + 1):
        sl = slice(i * sub_num, min((i + 1) * sub_num, n))
        if points[sl].shape[0] == 0:
            continue
        nn, gap = knn_gaps(tree, points[sl], k)
        nbr[sl], kg[sl] = nn, gap
        o = torch.from_numpy(cur[nn])
        mg[sl] = _medoid_gap(o)
        center, _ = O.compute_points_similarity(o)
        so = torch.from_numpy(cur[sl])
        sim = torch.maximum(torch.cosine_similarity(center, so, dim=-1), torch.cosine_similarity(center, -so, dim=-1))
        ug[sl] = np.abs(sim.double().numpy() - 0.95)
        cur[sl] = ori_out_oracle[sl]                      # what later chunks gather
    return dict(knn_gap=kg, medoid_gap=mg, update_gap=ug, nbr=nbr)


def taint_closure(ambiguous, nbr, sub_num=5000):
    """points whose result may legitimately differ: ambiguous themselves, or gathering a tainted point of an earlier chunk."""
    n = ambiguous.shape[0]
    t = ambiguous.copy()
    chunk = np.arange(n) // sub_num
    for i in range(n // sub_num + 1):
        sl = slice(i * sub_num, min((i + 1) * sub_num, n))
        if sl.start >= n:
            break
        nb = nbr[sl]
        earlier = chunk[nb] < i
        t[sl] |= np.any(t[nb] & earlier, axis=1)
    return t


def fuse_margins(select_points, select_ori, grid=(256, 256, 192), voxel_min=(-0.32, -0.32, -0.24), voxel_size=0.005 / 2):
    """voxelisation (PMVO.py:695-726): per point the distance of (p - min)/vsize from a rounding boundary (p2v, in voxel
    units), per occupied voxel the medoid gap.  -> (round_gap [n], voxel keys [m] (x*gy*gz + y*gz + z), medoid_gap [m])"""
    g = np.array(grid).astype(np.int64)
    p = np.array(select_points, dtype=np.float64, copy=True)
    p[:, 1:] *= -1
    q = (p - np.array(voxel_min)) / voxel_size
    rg = np.min(np.abs((q - np.floor(q)) - 0.5), axis=1)
    idx = np.clip(np.round(q).astype(np.int64), 0, g - 1)
    key = (idx[:, 0] * g[1] + idx[:, 1]) * g[2] + idx[:, 2]
    so = np.array(select_ori, copy=True)
    so[so[:, 1] > 0] *= -1
    order = np.argsort(key, kind="stable")
    ks = key[order]
    starts = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    ends = np.r_[starts[1:], ks.size]
    keys, gaps = [], []
    for s, e in zip(starts, ends):
        val = torch.from_numpy(so[order[s:e]]).type(torch.float)
        keys.append(ks[s])
        gaps.append(_medoid_gap(val[None])[0])
    return rg, np.array(keys, dtype=np.int64), np.array(gaps)
