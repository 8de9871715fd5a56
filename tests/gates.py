"""Margin-gated exact comparisons (SURVEY.md §7 "Hard parts"; margins: oracle/margins.py, stored in the goldens by
tests/golden/add_margins.py).  An item enters the strict comparison unless the ORACLE says its discrete decision was
closer to flipping than the eps below; excluded items are counted and printed, and may be at most a small share.

The kernels follow the reference's fp32 operation order, so the eps are a few ulps of the compared quantity, not
"percent of items allowed to differ":
  EPS_FWD   2e-8  winning vs runner-up (base view, depth sample) loss; a loss is sum(l*w)/sum(w) with sums of a few units,
                  and the one known order deviation (torch.sum's interleaved tail columns) moves either sum by
                  an ulp: <= ~1e-8 on a loss, so two candidate losses closer than 2e-8 may legitimately swap (1.2 % of the
                  items at BASELINE scale)
  EPS_THR   1e-7  | sum(w)/count - conf_threshold | of the confidence tests that feed that choice (same sums, same deviation)
  singleton       points the reference sampled in a batch of one: torch.matmul takes MKL's matrix-vector path there, whose
                  accumulation order differs from the sgemm path every other point sees, i.e. the reference's own result
                  for such a point depends on what else is in its forward() chunk (oracle/margins.py:singleton_base_groups;
                  0.5 % of the items at BASELINE scale).  The kernel reproduces the sgemm order; these points are listed
                  and compared at 5e-4 on the direction (an ulp or two of a coordinate over a sub-millimetre step) instead of bit
                  for bit.
  EPS_KNN   1e-13 gap between consecutive neighbour distances (float64, ~1e-3 m): set membership and summation order
  EPS_UPD   2e-7  | |cos(center, ori)| - 0.95 |, the update threshold of refine step (i)
  EPS_SEL   2e-8  | refine loss - threshold |, membership of the selected set
  EPS_ROUND 1e-9  distance of (p - min)/vsize from a .5 rounding boundary, in voxels (float64 index math)
Medoid means are reproduced bit for bit (torch.mean's order), so a medoid needs no margin once its neighbour order is
certain; the number of medoids whose top-2 gap is below 1e-6 is printed to show what that exactness buys.
"""
import numpy as np

EPS_FWD, EPS_THR, EPS_KNN, EPS_UPD, EPS_SEL, EPS_ROUND = 2e-8, 1e-7, 1e-13, 2e-7, 2e-8, 1e-9
MAX_EXCLUDED = 0.05


def forward_gate(g):
    """-> (keep, singleton): strict set, and the batch-of-one points taken out of it"""
    keep = g["fwd_margin"] > EPS_FWD
    if "fwd_thr_gap" in g:
        keep &= g["fwd_thr_gap"] > EPS_THR
    single = np.asarray(g["fwd_singleton"]) if "fwd_singleton" in g else np.zeros(len(keep), dtype=bool)
    return keep & ~single, single


def check_forward(g, ori, loss, hc, what="forward"):
    keep, single = forward_gate(g)
    n, ex = len(keep), int((~keep).sum())
    exact = np.all(ori == g["fwd_ori"], axis=1)
    dl = np.abs(loss.astype(np.float64) - g["fwd_loss"])
    print(f"\n{what}: {n} points, {ex} excluded (oracle margin <= {EPS_FWD:g} / threshold gap <= {EPS_THR:g}: {ex - int(single.sum())}, "
          f"batch-of-one in the reference: {int(single.sum())}); strict set: {exact[keep].mean() * 100:.3f}% "
          f"directions bit-identical, max|dloss| {dl[keep].max():.3g}; excluded set: {exact[~keep].mean() * 100 if ex else 100:.1f}% identical")
    sure = single & (g["fwd_margin"] > 1e-6)                       # batch-of-one points with a clear decision: same choice, ulp-level samples
    if sure.any():
        dev = np.abs(ori[sure] - g["fwd_ori"][sure]).max()
        print(f"  batch-of-one points with margin > 1e-6: {int(sure.sum())}, {int(exact[sure].sum())} bit-identical, max|d ori| {dev:.3g}")
        assert dev <= 5e-4                                      # 1-2 ulp of a ~1 m coordinate over a >= 0.5 mm step
    assert ex <= MAX_EXCLUDED * n, f"{ex} of {n} items excluded by the margin gate"
    for i in np.flatnonzero(~exact & keep)[:8]:
        print(f"  differs: point {i}  margin {g['fwd_margin'][i]:.3g}  loss ref {g['fwd_loss'][i]!r} gpu {loss[i]!r}  ori ref {g['fwd_ori'][i]} gpu {ori[i]}")
    assert exact[keep].all(), f"{int((~exact[keep]).sum())} directions differ on items with margin > {EPS_FWD:g}"
    assert dl[keep].max() <= 2e-8
    assert np.array_equal(hc[keep], g["fwd_hc"][keep])
    assert dl.max() <= 1e-5                                        # every item: the loss itself never moves by more


def refine_gate(g, sub_num=5000):
    """-> (clean [n] bool, selected-set certain bool)"""
    from oracle.margins import taint_closure
    amb = (g["ref_knn_gap"] <= EPS_KNN) | (g["ref_update_gap"] <= EPS_UPD)
    tainted = taint_closure(amb, g["ref_nbr"].astype(np.int64), sub_num)
    sel_certain = bool(np.all(np.abs(g["ref_min_loss"].astype(np.float64) - float(g["thr"])) > EPS_SEL))
    return ~tainted, sel_certain


def check_refine(g, so, ml, what="refine"):
    clean, sel_certain = refine_gate(g)
    n, ex = len(clean), int((~clean).sum())
    same = np.all(so == g["ref_select_o"], axis=1)
    dl = np.abs(ml.astype(np.float64) - g["ref_min_loss"])
    print(f"\n{what}: {n} points, {ex} excluded (kNN gap <= {EPS_KNN:g} / update gap <= {EPS_UPD:g}, closed over the chunk order); "
          f"strict set: {same[clean].mean() * 100:.3f}% rows bit-identical, max|dloss| {dl[clean].max():.3g}; "
          f"medoids with top-2 gap < 1e-6: {int((g['ref_medoid_gap'] < 1e-6).sum())}")
    assert ex <= MAX_EXCLUDED * n
    assert same[clean].all(), f"{int((~same[clean]).sum())} refined orientations differ on unambiguous points"
    assert dl[clean].max() <= EPS_SEL
    return clean, sel_certain


def check_near_surface(g, fu_p, fu_o, sel_certain):
    assert np.array_equal(fu_p, g["ref_fu_points"])                 # head filter decisions: exact
    if not sel_certain:
        print("near-surface orientations: selected set within eps of the threshold, comparison skipped")
        return
    # rows of ref_fu_* are the inputs that passed the head filter, in order
    keep_in = np.zeros(len(g["filter_unvisible_in"]), dtype=bool)
    j = 0
    for i, p in enumerate(g["filter_unvisible_in"].astype(np.float32)):
        if j < len(fu_p) and np.array_equal(p, fu_p[j]):
            keep_in[i] = True
            j += 1
    gap = g["fu_knn_gap"][keep_in] if len(g["fu_knn_gap"]) == len(keep_in) else np.full(len(fu_p), np.inf)
    ok = gap > EPS_KNN
    same = np.all(fu_o == g["ref_fu_ori"], axis=1)
    print(f"near-surface: {len(fu_p)} points, {int((~ok).sum())} excluded (kNN gap); strict set {same[ok].mean() * 100 if ok.any() else 100:.3f}% "
          f"bit-identical; medoids with top-2 gap < 1e-6: {int((g['fu_medoid_gap'] < 1e-6).sum())}")
    assert (~ok).sum() <= MAX_EXCLUDED * max(len(ok), 1)
    assert same[ok].all()


def check_volume(g, Occ, Ori, clean_all, sel_certain, what="orientation volume"):
    """Occ/Ori as loaded from the .mat pair.  clean_all: True, or bool over the fused points (selected + near-surface)."""
    nz = np.argwhere(Occ > 0)
    Z = Occ.shape[2]
    if not sel_certain or np.any(g["vox_round_gap"] <= EPS_ROUND):
        print(f"{what}: a point within eps of a selection / rounding boundary, occupancy comparison skipped")
        return
    assert np.array_equal(nz, g["mat_occ_nz"])                     # occupancy: bit-exact
    vals = np.stack([Ori[nz[:, 0], nz[:, 1], nz[:, 2] + c * Z] for c in range(3)], 1) if len(nz) else np.zeros((0, 3))
    same = np.all(vals == g["mat_ori_nz"], axis=1)
    print(f"{what}: {len(nz)} occupied voxels, {same.mean() * 100:.3f}% bit-identical; voxel medoids with top-2 gap < 1e-6: "
          f"{int((g['vox_medoid_gap'] < 1e-6).sum())}")
    if clean_all is True:
        assert same.all(), f"{int((~same).sum())} voxels differ"
    else:
        assert (~same).sum() <= (~clean_all).sum(), "more voxels differ than there are ambiguous points"
