"""GPU edge cases the reference's behaviour defines implicitly: degenerate points (NaN propagation), ragged sizes,
minimum / odd view counts, other patch sizes, trace termination rules."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pmvo(V, H, W, P, seed, conf_thr=0.15):
    from monohair_b200 import synthetic as syn
    from monohair_b200.camera import cameras_from_scene
    from monohair_b200.pmvo import PMVO
    from oracle import pmvo_oracle as O
    sc = syn.make_scene(V=V, H=H, W=W, seed=seed)
    Ori, Conf = sc.ref_ori_conf()
    pm = PMVO(cameras_from_scene(sc), sc.ref_depths(), Ori, Conf, sc.ref_masks(), device="cuda:0", image_size=[H, W],
              patch_size=P, visible_threshold=1, conf_threshold=conf_thr)
    return sc, pm, O.ViewMaps.from_scene(sc)


@pytest.mark.parametrize("V,H,W,P", [(20, 90, 150, 3), (21, 101, 163, 9), (33, 64, 64, 1), (23, 120, 90, 11)])
def test_forward_and_filter_vs_oracle_other_shapes(V, H, W, P):
    from monohair_b200 import synthetic as syn
    from oracle import pmvo_oracle as O
    sc, pm, vm = _pmvo(V, H, W, P, seed=V)
    cand = syn.candidate_points(n_cells=700, num_per_grid=2, seed=V)
    pts = torch.from_numpy(cand).float()
    so, fo, cnt_o = O.filter_points(vm, pts, P, 1, 0.15)
    s, sp, f = pm.filter_points(pts)
    assert np.array_equal(s.cpu().numpy(), so.numpy()) and np.array_equal(f.cpu().numpy(), fo.numpy())
    sel = cand[so.numpy()][:40]
    _, o_o, l_o, hc_o = O.forward(vm, sel, P, 0.15)
    _, ori, loss, hc = pm.forward(sel)
    assert np.abs(loss.cpu().numpy() - l_o.numpy()).max() <= 1e-5
    assert np.mean(np.abs(ori.cpu().numpy() - o_o.numpy()).max(1) <= 1e-4) >= 0.9
    assert np.array_equal(hc.cpu().numpy(), hc_o.numpy())


def test_degenerate_points_nan_semantics_match_oracle():
    """points no view sees: all weights are 0 -> 0/0 in the reference (PMVO.py:198-201); NaN must propagate the same way."""
    from oracle import pmvo_oracle as O
    sc, pm, vm = _pmvo(22, 96, 128, 5, seed=9)
    far = np.array([[3.0, 3.0, 3.0], [0.0, 0.0, 0.0], [0.0, 5.0, 0.0], [-0.101, 0.0, 0.0]])   # outside / centre of the head
    _, o_o, l_o, hc_o = O.forward(vm, far, 5, 0.15)
    _, ori, loss, hc = pm.forward(far)
    l, lo = loss.cpu().numpy(), l_o.numpy()
    assert np.array_equal(np.isnan(l), np.isnan(lo))
    assert np.allclose(l[~np.isnan(lo)], lo[~np.isnan(lo)], atol=1e-5)
    assert np.array_equal(hc.cpu().numpy(), hc_o.numpy())
    s, sp, f = pm.filter_points(torch.from_numpy(far).float())
    so, fo, _ = O.filter_points(vm, torch.from_numpy(far).float(), 5, 1, 0.15)
    assert np.array_equal(s.cpu().numpy(), so.numpy()) and np.array_equal(f.cpu().numpy(), fo.numpy())


def test_forward_requires_twenty_views():
    from monohair_b200._lib import MonoHairError
    sc, pm, vm = _pmvo(12, 64, 64, 3, seed=1)
    with pytest.raises(MonoHairError, match="20 views"):
        pm.forward(np.zeros((4, 3)))


def _solver_from(occ, ori):
    """occ [Z,Y,X], ori [3,Z,Y,X] in HairGrowing's frame -> solver + oracle volume."""
    from monohair_b200.hairgrow import HairGrowing
    from oracle import hairgrow_oracle as Hh
    vol = torch.zeros(occ.shape + (4,), device="cuda:0")
    vol[..., :3] = torch.from_numpy(ori).permute(1, 2, 3, 0).cuda()
    vol[..., 3] = torch.from_numpy(occ).cuda()
    return HairGrowing(volume=vol, device="cuda:0"), Hh.Volume(occ, ori)


def test_trace_termination_rules_vs_oracle():
    from oracle import hairgrow_oracle as Hh
    Z, Y, X = 24, 300, 20
    occ = np.zeros((Z, Y, X), np.float32)
    ori = np.zeros((3, Z, Y, X), np.float32)
    occ[10:14, :, 8:12] = 1                       # a long straight column along +y: hits the 256-step cap both ways
    ori[1, 10:14, :, 8:12] = 1.0
    ori[0, 10:14, 200:, 8:12] = 0.8               # a kink (dot = 0.78 < 0.85) stops the walk at y = 200
    ori[1, 10:14, 200:, 8:12] = 0.6
    occ[2:4, 5:9, 2:4] = 1                        # a blob with zero orientation: never moves, dot = 0 < thr -> < 5 points
    hg, volc = _solver_from(occ, ori)
    seeds = np.array([[9.2, 150.3, 11.7], [9.9, 20.1, 12.5], [9.5, 290.0, 10.2], [2.5, 6.5, 2.5], [-3.0, 150.0, 11.0],
                      [9.0, 400.0, 11.0], [15.0, 150.0, 11.0]], np.float32)
    pts, off, ln = hg._trace_batch(torch.from_numpy(seeds).cuda(), 0.85)
    flag = np.zeros_like(occ)
    for i, sd in enumerate(seeds):
        row = sd.copy() - 0.5                      # oracle.trace adds 0.5 + jitter*0.5: feed it the un-jittered seed
        s = Hh.trace(volc, row, flag, 0.85, np.zeros(3, np.float32))
        n = int(ln[i])
        if s is None:
            assert n == 0, (i, n)
        else:
            got = pts[int(off[i]):int(off[i]) + n].cpu().numpy()
            assert got.shape == s.shape and np.array_equal(got, s), i
    assert int(ln.max()) <= 513 and int(ln[0]) > 200


def test_trace_empty_volume_and_scalp_none():
    occ = np.zeros((8, 8, 8), np.float32)
    ori = np.zeros((3, 8, 8, 8), np.float32)
    hg, volc = _solver_from(occ, ori)
    strands = hg.randomlyGenerateSegments(0.85)
    assert strands == []
    roots = torch.tensor([[4.0, 1.0, 4.0]])
    normals = torch.tensor([[0.0, 1.0, 0.0]])
    s, num_root = hg.GenerateGuideStrandFromScalp(roots, normals, None, 0.85)
    assert s == [] and num_root == 0               # 25 inner steps through empty space -> None (HairGrow.py:216-221)


@pytest.mark.parametrize("H,W", [(37, 53), (128, 130), (9, 400)])
def test_gabor_ragged_sizes_vs_oracle(H, W):
    from monohair_b200.gabor import calOrientationGabor
    from oracle import gabor_oracle as G
    rng = np.random.default_rng(H * W)
    img = (rng.normal(0, 0.05, (H, W))).astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    img += (0.1 * np.cos(2 * np.pi * (xx * math.cos(0.7) + yy * math.sin(0.7)) / 4.3)).astype(np.float32)
    two_o, orient_o, conf_o, res = G.gabor_orientation(img)
    two, orient, conf = calOrientationGabor()(torch.from_numpy(img)[None, None].cuda())
    res = res.numpy()
    top2 = np.sort(res, axis=0)[-2:]
    ok = (top2[1] - top2[0]) / res.max() > 1e-4
    assert np.array_equal(orient[0, 0].cpu().numpy()[ok], orient_o.numpy()[ok])
    assert np.abs(conf[0, 0].cpu().numpy() - conf_o.numpy())[ok].max() <= 2e-3


def test_shared_reciprocal_division_is_ieee():
    """mh_div2 (one refined reciprocal for two quotients, mh_common.cuh) against div.rn over 2^28 operand triples:
    uniform bit patterns inside and outside its fast range, projection-like magnitudes, exact-quotient and
    near-tie cases, zeros / infinities / NaN / subnormals."""
    import ctypes as C
    from monohair_b200._lib import check, lib, ptr, stream_ptr
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(11)
    bad = torch.zeros((1,), dtype=torch.int64, device=dev)
    n = 1 << 24

    def run(a0, a1, b):
        check(lib().mh_debug_div2_check(stream_ptr(dev), ptr(a0.contiguous()), ptr(a1.contiguous()), ptr(b.contiguous()),
                                        a0.numel(), ptr(bad)), "mh_debug_div2_check")

    def rand_bits():
        return torch.randint(-2 ** 31, 2 ** 31 - 1, (n,), dtype=torch.int64, device=dev, generator=g).int().view(torch.float32)
    for _ in range(6):                                     # arbitrary bit patterns (mostly outside the fast range)
        run(rand_bits(), rand_bits(), rand_bits())
    for scale in (1.0, 1e-3, 1e3, 1e-9, 1e9, 5e-13, 2e12):  # log-uniform magnitudes around / across the range limits
        mk = lambda: (torch.exp((torch.rand(n, device=dev, generator=g) - 0.5) * 20) * scale *
                      torch.sign(torch.rand(n, device=dev, generator=g) - 0.5)).float()
        run(mk(), mk(), mk())
    for _ in range(3):                                     # projection-like: pixel-scale numerators over camera depth / offsets
        z = -(0.4 + torch.rand(n, device=dev, generator=g))
        run((torch.rand(n, device=dev, generator=g) - 0.5) * 3 * z, (torch.rand(n, device=dev, generator=g) - 0.5) * 3 * z, z)
        d = (torch.rand(n, device=dev, generator=g) - 0.5) * 40
        e = (torch.rand(n, device=dev, generator=g) - 0.5) * 40
        run(d, e, torch.sqrt(d * d + e * e))
    q = torch.randint(1, 1 << 23, (n,), device=dev, generator=g).float()      # exact quotients and their neighbours
    b = torch.randint(1, 1 << 12, (n,), device=dev, generator=g).float()
    run(q * b, torch.nextafter(q * b, torch.full_like(q, 1e30)), b)
    sp = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), float("nan"), 1e-45, -1e-40, 1.17549435e-38, 3.4e38, 1.0, -1.0,
                       9.094947e-13, 1.0995116e12], device=dev)
    a0, a1, bb = torch.meshgrid(sp, sp, sp, indexing="ij")
    run(a0.reshape(-1), a1.reshape(-1), bb.reshape(-1))
    torch.cuda.synchronize()
    assert int(bad.item()) == 0, f"{int(bad.item())} quotients differ from div.rn"


def test_candidate_sampling_on_device_equals_host_path():
    """SamplePointsAroundmesh (PMVO_utils.py:316-339): the device path (cell marking, np.nonzero-order compaction, float64
    sample arithmetic; random numbers from numpy's seeded global stream) is bit-identical to the numpy path, including
    points exactly on .5 cell boundaries and outside the grid."""
    from monohair_b200.pmvo_utils import SamplePointsAroundmesh
    rng = np.random.default_rng(3)
    pts = (rng.normal(size=(30000, 3)) * np.array([0.1, 0.13, 0.11]))
    bmin, vs = np.array([-0.32, -0.32, -0.24]), 0.005 / 4
    pts[:50] = np.round(pts[:50] / vs) * vs + vs / 2            # rounding boundaries
    pts[50:60] *= 10                                             # clipped
    np.random.seed(7)
    a = SamplePointsAroundmesh(pts.copy(), bmin, vs, num_per_grid=4, grid_resolution=[512, 512, 384])
    np.random.seed(7)
    p2 = pts.copy()
    b = SamplePointsAroundmesh(p2, bmin, vs, num_per_grid=4, grid_resolution=[512, 512, 384], device="cuda:0")
    assert a.shape == b.shape and a.shape[0] > 50000 and np.array_equal(a, b)
    assert np.array_equal(p2[:, 1:], -pts[:, 1:])                # the reference's in-place flip of the argument


def test_depth_rasteriser_vs_oracle_and_pmvo_convention(tmp_path):
    """mh_render_depth against the numpy restatement of the reference's OpenGL depth pass (oracle/render_oracle.py, parity
    unpinned: moderngl is absent), and against PMVO's own projection: a surface point projected by the PMVO kernels must see
    itself in the rendered depth map (visibility ~1), which ties the rasteriser's pixel convention to project_points."""
    from monohair_b200 import synthetic as syn
    from monohair_b200.camera import cameras_from_scene
    from monohair_b200.render import DepthRenderer
    from oracle import render_oracle as RO
    # an ellipsoid mesh (the synthetic capture's shape)
    nu, nv = 48, 96
    th, ph = np.meshgrid(np.linspace(0.05, np.pi - 0.05, nu), np.linspace(0, 2 * np.pi, nv, endpoint=False), indexing="ij")
    r = np.array(syn.RADII)
    V = np.stack([r[0] * np.sin(th) * np.cos(ph), r[1] * np.cos(th), r[2] * np.sin(th) * np.sin(ph)], -1).reshape(-1, 3)
    F = []
    for i in range(nu - 1):
        for j in range(nv):
            a, b, c, d = i * nv + j, i * nv + (j + 1) % nv, (i + 1) * nv + j, (i + 1) * nv + (j + 1) % nv
            F += [[a, c, b], [b, c, d]]
    F = np.array(F, np.int32)
    sc = syn.make_scene(V=20, H=90, W=160, seed=4)
    cams = cameras_from_scene(sc)
    R = DepthRenderer(sc.H, sc.W)
    R.add_mesh(V, F)
    for name in list(cams)[:3]:
        c = cams[name]
        d = R.draw(c).cpu().numpy()
        o = RO.render_depth(V, F, c.pose.numpy(), c.ndc_prj, sc.H, sc.W)
        cover_same = ((d < 1) == (o < 1))
        both = (d < 1) & (o < 1)
        assert cover_same.mean() > 0.999 and both.sum() > 500
        assert np.abs(d - o)[both].max() < 2e-5 or np.mean(np.abs(d - o)[both] < 2e-5) > 0.995   # silhouette-edge pixels may pick the other face
    # PMVO sees the rendered surface as visible: depth maps from the rasteriser, analytic everything else
    from monohair_b200.pmvo import PMVO
    depth = np.stack([R.draw(cams[k]).cpu().numpy() for k in cams]).astype(np.float32) * 255.0
    pm = PMVO.from_u8(cams, depth, sc.ori_gray, sc.conf_u8, sc.mask_u8, device="cuda:0", image_size=[sc.H, sc.W], patch_size=5,
                      visible_threshold=1, conf_threshold=0.15)
    pm.Compute_Visible_and_Ori(V[::7].astype(np.float32))
    vis = pm.visible.cpu().numpy()
    front = vis > 0.5
    assert front.sum() > 0.25 * vis.size                            # a surface point is visible from the cameras facing it
    assert np.median(vis[front]) > 0.9
