"""GPU parity tests of the PMVO path: CUDA kernels (through the C ABI) against
  (a) golden vectors produced by the unmodified reference (tests/golden/*.npz) and
  (b) the CPU oracle (oracle/pmvo_oracle.py) on the same seeded inputs.

Bars: pixel / voxel indices, visibility values, surface / filter masks, base-view selection: bit-exact.
Floats: min_loss <= 1e-5 abs, direction <= 1e-4 L-inf -- the kernels follow the reference's fp32
operation order, so in practice they are bit-identical for almost every point; the exact-match rate is printed.
"""
import numpy as np
import pytest
import torch
from scipy.spatial import KDTree

import gates
from golden_util import load, scene_of

pytestmark = pytest.mark.gpu

CASES = ["pmvo_p7", "pmvo_p5_ties"]
LOSS_ATOL = 1e-5
ORI_LINF = 1e-4


def build(g, sc):
    from monohair_b200.camera import cameras_from_scene
    from monohair_b200.pmvo import PMVO
    Ori, Conf = sc.ref_ori_conf()
    return PMVO(cameras_from_scene(sc), sc.ref_depths(), Ori, Conf, sc.ref_masks(), device="cuda:0",
                image_size=[sc.H, sc.W], patch_size=int(g["patch"]), visible_threshold=1,
                conf_threshold=float(g["conf_thr"]))


@pytest.fixture(scope="module", params=CASES)
def case(request):
    from monohair_b200 import pmvo as P
    g = load(request.param)
    sc = scene_of(g)
    P.scalp_tree, P.scalp_max = KDTree(data=g["scalp"]), g["scalp"].max(0)     # module globals of the reference (PMVO.py:99-106)
    return g, sc, build(g, sc)


def test_filter_points_vs_reference_golden(case):
    g, sc, pmvo = case
    pts = torch.from_numpy(g["points"][: int(g["n_covered"])]).float()
    s, sp, f = pmvo.filter_points(pts)
    assert np.array_equal(s.cpu().numpy(), g["surface_index"])
    assert np.array_equal(f.cpu().numpy(), g["filter_index"])
    assert np.array_equal(sp.cpu().numpy(), pts.numpy()[g["surface_index"]])


def test_filter_counters_and_centre_values_bit_exact_vs_oracle(case):
    from oracle import pmvo_oracle as O
    g, sc, pmvo = case
    vm = O.ViewMaps.from_scene(sc)
    pts = torch.from_numpy(g["points"][:3000]).float()
    _, _, cnt_o = O.filter_points(vm, pts, int(g["patch"]), 1, float(g["conf_thr"]))
    _, cnt = pmvo.filter_counters(pts)
    assert torch.equal(cnt.cpu(), cnt_o)
    st = O.compute_visible_and_ori(vm, pts[:500], int(g["patch"]))
    pmvo.Compute_Visible_and_Ori(pts[:500])
    assert torch.equal(pmvo.visible.cpu(), st["visible"])
    assert torch.equal(pmvo.Conf.cpu(), st["Conf"])
    assert torch.equal(pmvo.Ori.cpu(), st["Ori"])
    unv = pmvo.compute_unvisible_points(pts)
    assert torch.equal(unv.cpu(), O.compute_unvisible_points(vm, pts))


def test_packed_via_u8_path_identical(case):
    from monohair_b200.camera import cameras_from_scene
    from monohair_b200.pmvo import PMVO
    g, sc, pmvo = case
    p2 = PMVO.from_u8(cameras_from_scene(sc), sc.depth, sc.ori_gray, sc.conf_u8, sc.mask_u8, device="cuda:0",
                      image_size=[sc.H, sc.W], patch_size=int(g["patch"]), visible_threshold=1,
                      conf_threshold=float(g["conf_thr"]))
    assert torch.equal(p2.mapC, pmvo.mapC)
    assert torch.equal(p2.mapP, pmvo.mapP)


def test_forward_vs_reference_golden(case):
    g, sc, pmvo = case
    _, ori, loss, hc, dbg = pmvo.forward(g["fwd_points"], debug=True)
    ori, loss, hc = ori.cpu().numpy(), loss.cpu().numpy(), hc.cpu().numpy()
    # base views: same order as torch.topk on CPU (ties included)
    assert np.array_equal(dbg["base_val"].cpu().numpy(), g["fwd_base_val"])
    assert np.array_equal(dbg["base_idx"].cpu().numpy().astype(np.int64), g["fwd_base_idx"])
    # exact agreement on every point whose decision margin in the oracle exceeds eps; excluded count printed
    gates.check_forward(g, ori, loss, hc)


def test_forward_per_base_losses_vs_oracle(case):
    from oracle import pmvo_oracle as O
    g, sc, pmvo = case
    vm = O.ViewMaps.from_scene(sc)
    n = 64
    _, o_o, l_o, hc_o, dbg_o = O.forward(vm, g["fwd_points"][:n], int(g["patch"]), float(g["conf_thr"]), debug=True)
    _, ori, loss, hc, dbg = pmvo.forward(g["fwd_points"][:n], debug=True)
    lb = dbg["loss_b"].cpu().numpy()
    ab = dbg["arg_b"].cpu().numpy()
    valid = ab >= 0
    lo = torch.stack(dbg_o["loss_b"]).numpy()
    ao = torch.stack(dbg_o["arg"]).numpy()
    # intermediate per-base minima (most are discarded by the best-over-bases selection): an ulp-level difference
    # of a projected direction can pick another patch entry for one (view, sample) and shift that view's weight
    dlb = np.abs(lb[valid] - lo[valid])
    assert np.mean(dlb <= LOSS_ATOL) >= 0.99 and dlb.max() <= 1e-3, (np.mean(dlb <= LOSS_ATOL), dlb.max())
    print(f"\nper-base: losses bit-identical {np.mean(lb[valid] == lo[valid]) * 100:.2f}%, argmin identical "
          f"{np.mean(ab[valid] == ao[valid]) * 100:.2f}%")
    assert np.mean(ab[valid] == ao[valid]) >= 0.97
    # bases the reference ignores (base conf <= 0, b > 0) are skipped by the kernel
    bc = dbg_o["base_conf"].numpy()
    assert np.array_equal(valid[1:], bc[1:] > 0)


def test_refine_voxelise_vs_reference_golden(case, tmp_path):
    import types
    import scipy.io
    from monohair_b200 import pmvo as P
    g, sc, pmvo = case
    scalp = g["scalp"]
    P.scalp_tree = KDTree(data=scalp)
    P.scalp_max = scalp.max(0)
    td = str(tmp_path)
    a = types.SimpleNamespace(output_path=td, save_path=td + "/refine", device="cuda:0",
                              PMVO=types.SimpleNamespace(visible_threshold=1), data=types.SimpleNamespace(root=td))
    import os
    os.makedirs(a.save_path, exist_ok=True)
    P.refine(g["fwd_points"].astype(np.float32), g["fwd_ori"].copy(), g["fwd_loss"].copy(), pmvo,
             g["filter_unvisible_in"].copy(), a, infer_inner=False, threshold=float(g["thr"]), genrate_ori_only=False)
    so = np.load(td + "/refine/select_o.npy")
    ml = np.load(td + "/refine/min_loss.npy")
    clean, sel_certain = gates.check_refine(g, so, ml)
    fu_p = np.load(td + "/refine/filter_unvisible.npy")
    fu_o = np.load(td + "/refine/filter_unvisible_ori.npy")
    gates.check_near_surface(g, fu_p, fu_o, sel_certain)
    Occ = scipy.io.loadmat(td + "/refine/Occ3D.mat")["Occ"]
    Ori = scipy.io.loadmat(td + "/refine/Ori3D.mat")["Ori"]
    assert Occ.dtype == np.float64 and Ori.dtype == np.float64
    assert tuple(Ori.shape) == tuple(g["mat_ori_shape"])
    Z = Occ.shape[2]
    gates.check_volume(g, Occ, Ori, True if clean.all() else np.concatenate([clean[ml < float(g["thr"])], np.ones(len(fu_p), bool)]),
                       sel_certain)
    assert np.all(Ori.reshape(Occ.shape[0], Occ.shape[1], 3, Z).transpose(0, 1, 3, 2)[Occ == 0] == 0)


def test_refine_chunks_one_call_equals_staged_pipeline(case):
    """mh_refine_chunks (sweep + re-score + finish in one C call) == pipeline.refine_stage (the three steps called
    separately, which is what the multi-GPU host shards); small chunks so that many chunk hand-overs happen."""
    from monohair_b200 import pipeline
    from monohair_b200 import pmvo as P
    from monohair_b200._lib import check, lib, ptr, stream_ptr
    g, sc, pmvo = case
    dev = pmvo.device
    pts = torch.from_numpy(g["fwd_points"].astype(np.float32)).to(dev).contiguous()
    ori = torch.from_numpy(g["fwd_ori"].astype(np.float32)).to(dev).contiguous()
    loss = torch.from_numpy(g["fwd_loss"].astype(np.float32)).to(dev).contiguous()
    n, k, sub = pts.size(0), 100, 37
    o_ref, l_ref = pipeline.refine_stage(pmvo, pts, ori, loss, sub_num=sub, k=k)
    nbr = P.knn(pts, pts, k, dev)
    filt = pmvo.filter_head_points(pts, pmvo.visible_threshold).to(torch.uint8).contiguous()
    o, l = ori.clone(), loss.clone()
    wsb = lib().mh_refine_chunks_workspace_bytes(n, sub)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    check(lib().mh_refine_chunks(stream_ptr(dev), pmvo._vp(), ptr(pts), ptr(nbr), k, ptr(filt), n, sub,
                                 float(pmvo.conf_threshold), ptr(o), ptr(l), ptr(ws), wsb), "mh_refine_chunks")
    assert torch.equal(o, o_ref) and torch.equal(l, l_ref)
    # and the sweep really is sequential across chunks: one big chunk (pure Jacobi) gives a different result
    o_j, _ = pipeline.refine_stage(pmvo, pts, ori, loss, sub_num=n, k=k)
    assert not torch.equal(o_j, o_ref)


@pytest.mark.parametrize("world", [1, 2, 3])
def test_distributed_sweep_ranks_simulated_on_one_gpu(case, world):
    """mh_refine_sweep_dist with `world` ranks played by `world` kernels on separate streams of ONE device (grids capped so
    that they are resident together), every rank with its own full copy of the output arrays: the cross-rank waits, the
    fan-out stores and the ownership formula run for real, and every copy must equal mh_refine_sweep's result bit for bit."""
    import ctypes as C
    from monohair_b200 import pipeline
    from monohair_b200 import pmvo as P
    from monohair_b200._lib import check, lib, ptr
    g, sc, pmvo = case
    dev = pmvo.device
    pts = torch.from_numpy(g["fwd_points"].astype(np.float32)).to(dev).contiguous()
    ori = torch.from_numpy(g["fwd_ori"].astype(np.float32)).to(dev).contiguous()
    n, k, sub = pts.size(0), 100, 37
    L = lib()
    nbr = P.knn(pts, pts, k, dev)
    o_ref, c_ref = torch.empty_like(ori), torch.empty_like(ori)
    wsb = L.mh_refine_sweep_workspace_bytes(n, sub)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    from monohair_b200._lib import stream_ptr
    check(L.mh_refine_sweep(stream_ptr(dev), ptr(ori), ptr(nbr), k, n, sub, ptr(o_ref), ptr(c_ref), ptr(ws), wsb), "mh_refine_sweep")
    bufs = [torch.empty((2, n, 3), dtype=torch.float32, device=dev) for _ in range(world)]
    for b in bufs:
        b[0].view(torch.int32).fill_(-1)                       # PENDING
        b[1].zero_()
    po = (C.c_uint64 * world)(*[b[0].data_ptr() for b in bufs])
    pc = (C.c_uint64 * world)(*[b[1].data_ptr() for b in bufs])
    B = int(L.mh_refine_sweep_dist_block())
    locals_ = [nbr[pipeline.block_cyclic_index(n, r, world, B, dev)].contiguous() for r in range(world)]
    errs = [torch.zeros((1,), dtype=torch.int32, device=dev) for _ in range(world)]
    scr = [torch.empty((64,), dtype=torch.uint8, device=dev) for _ in range(world)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    torch.cuda.synchronize()
    for r in reversed(range(world)):                          # launch the LAST rank first: rank 0's points must still get through
        check(L.mh_refine_sweep_dist(C.c_void_p(streams[r].cuda_stream), ptr(ori), ptr(locals_[r]), k, n, sub, r, world, po, pc,
                                     5.0, 48, ptr(scr[r]), 64, ptr(errs[r])), "mh_refine_sweep_dist")
    torch.cuda.synchronize()
    assert all(int(e.item()) == 0 for e in errs), "a wait ran out"
    for b in bufs:
        assert torch.equal(b[0], o_ref) and torch.equal(b[1], c_ref)


def test_voxel_fuse_vs_oracle_exact(case):
    from oracle import pmvo_oracle as O
    from monohair_b200 import pmvo as P
    g, sc, pmvo = case
    rng = np.random.default_rng(5)
    n = 6000
    pts = (g["fwd_points"][rng.integers(0, len(g["fwd_points"]), n)] + rng.normal(0, 2e-3, (n, 3))).astype(np.float32)
    dirs = rng.normal(size=(n, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    # exact .5 boundaries of the float64 index math (np.round half-even)
    pts[:50, 0] = (-0.32 + (np.arange(50) + 100.5) * 0.0025).astype(np.float32)
    occ_o, ori_o = O.voxel_fuse(pts.copy(), dirs.copy())
    vol, vidx = P.voxel_fuse(pts, dirs, "cuda:0", return_index=True)
    x, y, z = O.p2v(pts.astype(np.float64).copy() * 1.0, np.array([-0.32, -0.32, -0.24]), 0.005 / 2, np.array([256, 256, 192]))
    assert np.array_equal(vidx.cpu().numpy(), (x.astype(np.int64) * 256 + y) * 192 + z)
    v = vol.cpu().numpy()                                              # [gz,gy,gx,4]
    occ = v[..., 3].transpose(2, 1, 0)
    assert np.array_equal(occ, occ_o.astype(np.float32))
    ori = np.stack([v[..., 0], -v[..., 1], -v[..., 2]], -1).transpose(2, 1, 0, 3)
    same = np.all(ori == ori_o.astype(np.float32), axis=-1)
    assert same[occ_o > 0].mean() >= 0.995
    assert same[occ_o == 0].all()
    om, orim = P.volume_to_mat(vol)
    mo, mori = O.mat_layout(occ.astype(np.float64), ori.astype(np.float64))
    assert np.array_equal(om.cpu().numpy(), mo) and np.array_equal(orim.cpu().numpy(), mori)


def test_voxel_fuse_crowded_voxels_vs_oracle():
    """Voxels holding 1 .. ~150 points: records of more than 32 entries take the overflow chain + the big-record
    kernel; small ones (K < 8) torch's scalar summation path.  Volume equals the oracle's; the persistent plane is
    left all-zero."""
    from oracle import pmvo_oracle as O
    from monohair_b200 import pmvo as P
    rng = np.random.default_rng(11)
    centres = rng.uniform([-0.2, -0.2, -0.15], [0.2, 0.2, 0.15], (400, 3))
    counts = rng.integers(1, 150, 400)
    counts[:60] = rng.integers(1, 9, 60)
    pts = np.concatenate([c + rng.uniform(-0.0012, 0.0012, (k, 3)) for c, k in zip(centres, counts)]).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]
    base = rng.normal(size=(1, 3))
    dirs = (base + 0.3 * rng.normal(size=(len(pts), 3))).astype(np.float32)
    occ_o, ori_o = O.voxel_fuse(pts.copy(), dirs.copy())
    for _ in range(2):                                                 # twice: the second call relies on the clean plane
        vol = P.voxel_fuse(pts, dirs, "cuda:0")
    v = vol.cpu().numpy()
    occ = v[..., 3].transpose(2, 1, 0)
    assert np.array_equal(occ, occ_o.astype(np.float32))
    ori = np.stack([v[..., 0], -v[..., 1], -v[..., 2]], -1).transpose(2, 1, 0, 3)
    same = np.all(ori == ori_o.astype(np.float32), axis=-1)
    print(f"crowded voxels: {same[occ_o > 0].mean() * 100:.2f}% of {int((occ_o > 0).sum())} voxels bit-identical")
    assert same[occ_o > 0].mean() >= 0.99
    assert same[occ_o == 0].all()
    _, plane = P.fuse_plane(torch.device("cuda:0"), P.GRID)
    assert int(plane[P.FUSE_HDR_BYTES:].view(torch.int64).abs().max().item()) == 0


def test_knn_exact_vs_kdtree():
    from monohair_b200 import pmvo as P
    rng = np.random.default_rng(0)
    ref = (rng.normal(size=(20000, 3)) * np.array([0.1, 0.13, 0.11])).astype(np.float32)
    ref /= np.maximum(np.linalg.norm(ref / np.array([0.1, 0.13, 0.11], np.float32), axis=1, keepdims=True), 1e-6)
    ref += rng.normal(0, 2e-3, ref.shape).astype(np.float32)
    q = np.concatenate([ref[:3000], (ref[:500] + 0.01).astype(np.float32), np.array([[1, 1, 1], [-1, 0, 0]], np.float32)])
    d_ref, i_ref = KDTree(data=ref).query(q, 100)
    idx = P.knn(torch.from_numpy(ref).cuda(), torch.from_numpy(q).cuda(), 100, torch.device("cuda:0")).cpu().numpy()
    dd = np.linalg.norm(ref[idx].astype(np.float64) - q[:, None, :].astype(np.float64), axis=-1)
    assert np.allclose(dd, d_ref, rtol=0, atol=1e-12)
    assert np.mean(idx == i_ref) > 0.9999


def test_knn_float64_queries_match_kdtree_where_float32_queries_do_not():
    """PMVO.refine step (iii) queries the KDTree with the float64 candidates and casts them to float32 afterwards
    (PMVO.py:670-671): mh_knn_q64 must order the neighbours by the float64-query distances.  Queries are placed next to
    the bisector of two reference points, so that rounding the query to float32 flips the order for some of them."""
    from monohair_b200 import pmvo as P
    rng = np.random.default_rng(1)
    ref = rng.uniform(-0.1, 0.1, (20000, 3)).astype(np.float32)
    tree = KDTree(data=ref)
    seeds = rng.uniform(-0.08, 0.08, (3000, 3))
    _, nn = tree.query(seeds, 100)
    a, b = ref[nn[:, 98]].astype(np.float64), ref[nn[:, 99]].astype(np.float64)
    q = 0.5 * (a + b) + rng.normal(0, 1e-9, a.shape)               # float64, ~1e-9 off the bisector of two close-ranked points
    d64, i64 = tree.query(q, 100)
    _, i32 = tree.query(q.astype(np.float32), 100)
    dev = torch.device("cuda:0")
    refd = torch.from_numpy(ref).to(dev)
    got64 = P.knn(refd, torch.from_numpy(q).to(dev), 100, dev).cpu().numpy()
    got32 = P.knn(refd, torch.from_numpy(q.astype(np.float32)).to(dev), 100, dev).cpu().numpy()
    flips = int((i64 != i32).any(1).sum())
    print(f"\nfloat32-rounded queries change the neighbour list of {flips} of {len(q)} queries")
    assert flips > 20                                              # the case exists
    dd = np.linalg.norm(ref[got64].astype(np.float64) - q[:, None, :], axis=-1)
    assert np.allclose(dd, d64, rtol=0, atol=1e-13)
    print(f"entries equal to scipy: float64 queries {np.mean(got64 == i64):.6f}, float32 queries {np.mean(got32 == i32):.6f}; "
          f"lists equal to the float64 KDTree answer: {np.mean((got64 == i64).all(1)):.4f} (float64 queries) vs "
          f"{np.mean((got32 == i64).all(1)):.4f} (float32 queries)")
    assert np.mean(got64 == i64) > 0.9999
    assert np.mean(got32 == i32) > 0.999          # rounded queries sit within 1e-16 of exact ties: order is implementation-defined
    assert np.mean((got64 == i64).all(1)) > 0.999 > np.mean((got32 == i64).all(1))


def test_knn_fallback_paths():
    """dense clumps (buffer overflow) and hundreds of coincident points (no separating radius) take the general
    kernel; results must still be the exact k nearest by distance."""
    from monohair_b200 import pmvo as P
    rng = np.random.default_rng(3)
    base = rng.uniform(-0.1, 0.1, (4000, 3)).astype(np.float32)
    clump = (np.array([[0.01, 0.02, 0.03]], np.float32) + rng.normal(0, 1e-5, (3000, 3))).astype(np.float32)
    dup = np.repeat(np.array([[-0.05, 0.0, 0.05]], np.float32), 400, axis=0)
    ref = np.concatenate([base, clump, dup])
    q = np.concatenate([ref[::7], np.array([[0.01, 0.02, 0.03], [-0.05, 0.0, 0.05]], np.float32)])
    d_ref, _ = KDTree(data=ref).query(q, 100)
    idx = P.knn(torch.from_numpy(ref).cuda(), torch.from_numpy(q).cuda(), 100, torch.device("cuda:0")).cpu().numpy()
    dd = np.linalg.norm(ref[idx].astype(np.float64) - q[:, None, :].astype(np.float64), axis=-1)
    assert np.allclose(dd, d_ref, rtol=0, atol=1e-12)
    assert all(len(set(r)) == 100 for r in idx[::50])


def test_empty_and_single_inputs(case):
    g, sc, pmvo = case
    s, sp, f = pmvo.filter_points(torch.zeros((0, 3)))
    assert s.numel() == 0 and sp.shape == (0, 3) and f.numel() == 0
    p, o, l, hc = pmvo.forward(np.zeros((0, 3)))
    assert o.shape == (0, 3) and l.numel() == 0
    p, o, l, hc = pmvo.forward(g["fwd_points"][:1])
    assert np.abs(l.cpu().numpy() - g["fwd_loss"][:1]).max() <= LOSS_ATOL
    far = np.array([[5.0, 5.0, 5.0]])                                   # out of every image: invisible everywhere
    s, sp, f = pmvo.filter_points(torch.from_numpy(far).float())
    assert not bool(s[0]) and not bool(f[0])


def test_voxel_fuse_winner_exchange_equals_dense_fusion():
    """the multi-GPU exchange format: disjoint voxel slabs fused to winner lists, concatenated and scattered, give the
    single fusion's volume bit for bit (what pipeline.fuse_stage does across ranks)."""
    from monohair_b200 import pmvo as P
    rng = np.random.default_rng(5)
    n = 60000
    pts = (rng.normal(size=(n, 3)) * np.array([0.05, 0.06, 0.05])).astype(np.float32)
    dirs = rng.normal(size=(n, 3)).astype(np.float32)
    dev = torch.device("cuda:0")
    tp, td = torch.from_numpy(pts).to(dev), torch.from_numpy(dirs).to(dev)
    vol = P.voxel_fuse(tp, td, dev)
    z = torch.round((-(tp[:, 2].double()) - float(P.VOXEL_MIN[2])) / float(P.VOXEL_SIZE)).clamp_(0, P.GRID[2] - 1).long()
    parts = []
    for a, b in ((0, 90), (90, 100), (100, 192)):
        win, cnt = P.voxel_fuse_winners(tp, td, dev, valid=(z >= a) & (z < b))
        c = int(cnt.item())
        assert bool((win[c:, 3].view(torch.int32) == -1).all()) and bool((win[:c, 3].view(torch.int32) >= 0).all())
        parts.append(win)                                  # padded entries (key -1) are skipped by the scatter
    vol2 = P.voxel_scatter(torch.cat(parts, 0), dev)
    assert torch.equal(vol, vol2)
    assert int(vol[..., 3].sum().item()) > 1000
