"""GPU, BASELINE.json configs[1] size (60 views at 1920x1080, patch 7, ~2.0 M candidates, grid 256x256x192):
size-independent properties of the PMVO path plus an oracle spot check on a random sub-sample at full resolution."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(V=60, H=1920, W=1080, patch=7, conf_thr=0.15, thr=0.025)


@pytest.fixture(scope="module")
def world():
    from monohair_b200 import synthetic as syn
    from monohair_b200 import pmvo as P
    from monohair_b200.camera import cameras_from_scene
    sc = syn.make_scene(V=CFG["V"], H=CFG["H"], W=CFG["W"], seed=0, device="cuda:0")
    cand = syn.candidate_points(num_per_grid=4, seed=0)
    scalp = syn.scalp_vertices(2000, seed=0)
    P.scalp_tree, P.scalp_max = scalp, scalp.max(0)
    pm = P.PMVO.from_u8(cameras_from_scene(sc), sc.depth, sc.ori_gray, sc.conf_u8, sc.mask_u8, device="cuda:0",
                        image_size=[sc.H, sc.W], patch_size=CFG["patch"], visible_threshold=1, conf_threshold=CFG["conf_thr"])
    return sc, cand, pm


def test_filter_properties_full_size(world):
    sc, cand, pm = world
    pts, cnt = pm.filter_counters(torch.from_numpy(cand).float())
    s, f = pm.filter_decide(cnt)
    c = cnt.cpu().numpy()
    assert cand.shape[0] > 1_900_000
    assert (c[0] <= CFG["V"]).all() and (c[3] >= c[0]).all()          # vis1 (looser depth test) dominates vis
    assert (c[1] <= c[0] + 1e-3).all() and (c[2] <= c[0]).all()       # masked / low-conf counts are sub-counts
    assert np.all(c[0] == np.round(c[0])) and np.all(c[3] == np.round(c[3]))
    assert not bool((s & f).any())                                    # surface and near-surface sets are disjoint
    assert 0.3 < float(s.float().mean()) < 0.7
    # determinism and batch invariance (points are independent)
    _, cnt2 = pm.filter_counters(torch.from_numpy(cand[100000:300000]).float())
    assert torch.equal(cnt2, cnt[:, 100000:300000])


def test_forward_properties_full_size(world):
    sc, cand, pm = world
    s, sp, f = pm.filter_points(torch.from_numpy(cand[:600000]).float())
    pts = sp[:60000].contiguous()
    _, ori, loss, hc = pm.forward(pts)
    assert torch.allclose(ori.norm(dim=1), torch.ones_like(loss), atol=1e-5)
    assert float(loss.min()) > -1e-5 and float(loss.max()) <= 1.0 + 1e-6
    # permutation / chunk invariance, determinism: bit-identical
    perm = torch.randperm(pts.size(0), device=pts.device, generator=torch.Generator(device=pts.device).manual_seed(0))
    _, o2, l2, h2 = pm.forward(pts[perm].contiguous())
    assert torch.equal(o2, ori[perm]) and torch.equal(l2, loss[perm]) and torch.equal(h2, hc[perm])
    _, o3, l3, _ = pm.forward(pts[:777].contiguous())
    assert torch.equal(o3, ori[:777]) and torch.equal(l3, loss[:777])
    # the recovered direction follows the synthetic flow field (sign-free), as it does for the reference
    from monohair_b200 import synthetic as syn
    T = syn.flow_tangent(pts.double(), syn.RADII).float()
    cosang = (T * ori).sum(1).abs()
    assert float(cosang.median()) > 0.99


def test_forward_oracle_spot_check_full_resolution(world):
    """64 random surface points through the CPU oracle with all 60 full-resolution views."""
    from oracle import pmvo_oracle as O
    sc, cand, pm = world
    s, sp, f = pm.filter_points(torch.from_numpy(cand[:900000]).float())
    g = torch.Generator().manual_seed(1)
    sel = sp.cpu()[torch.randperm(sp.size(0), generator=g)[:64]].numpy().astype(np.float64)
    vm = O.ViewMaps.from_scene(sc)
    _, o_o, l_o, hc_o = O.forward(vm, sel, CFG["patch"], CFG["conf_thr"])
    _, ori, loss, hc = pm.forward(sel)
    dl = np.abs(loss.cpu().numpy() - l_o.numpy()).max()
    do = np.abs(ori.cpu().numpy() - o_o.numpy()).max(axis=1)
    print(f"\nfull-res spot check: max |dloss| = {dl:.3g}; directions identical {np.mean(do == 0) * 100:.1f}%, "
          f"within 1e-4: {np.mean(do <= 1e-4) * 100:.1f}%")
    assert dl <= 1e-5
    assert np.mean(do <= 1e-4) >= 0.95
    assert np.array_equal(hc.cpu().numpy(), hc_o.numpy())
    # filter masks at full resolution, bit-exact
    so, fo, _ = O.filter_points(vm, torch.from_numpy(cand[:20000]).float(), CFG["patch"], 1, CFG["conf_thr"])
    assert np.array_equal(s[:20000].cpu().numpy(), so.numpy()) and np.array_equal(f[:20000].cpu().numpy(), fo.numpy())


def test_voxel_fusion_properties_full_size(world):
    from monohair_b200 import pmvo as P
    from monohair_b200 import synthetic as syn
    sc, cand, pm = world
    rng = np.random.default_rng(2)
    pts = torch.from_numpy(cand[rng.random(cand.shape[0]) < 0.8]).float().cuda()
    dirs = syn.flow_tangent(pts.double(), syn.RADII).float()
    dirs = dirs * torch.where(torch.rand(dirs.size(0), 1, device="cuda") < 0.5, -1.0, 1.0)     # random signs
    vol, vidx = P.voxel_fuse(pts, dirs, "cuda:0", return_index=True)
    gx, gy, gz = P.GRID
    occ = vol[..., 3]
    # occupancy == set of voxel keys, exactly
    key = vidx.long()
    x, rem = key // (gy * gz), key % (gy * gz)
    y, z = rem // gz, rem % gz
    ref = torch.zeros_like(occ)
    ref[z, y, x] = 1.0
    assert torch.equal(occ, ref)
    assert int(occ.sum().item()) == int(torch.unique(key).numel())
    o = vol[..., :3][occ > 0]
    assert torch.allclose(o.norm(dim=1), torch.ones(o.size(0), device="cuda"), atol=1e-5)
    assert bool((o[:, 1] >= 0).all())                      # stored as -ori.y with ori.y <= 0 (PMVO.py:702-703, HairGrow.py:55)
    assert bool((vol[..., :3][occ == 0] == 0).all())
    # the medoid is one of the voxel's own (flipped) input directions
    flipped = torch.where(dirs[:, 1:2] > 0, -dirs, dirs) * torch.tensor([1.0, -1.0, -1.0], device="cuda")
    got = vol[..., :3][z, y, x]                            # voxel value seen by every point
    dist = (flipped - got).abs().amax(1)
    first = torch.full((gx * gy * gz,), float("inf"), device="cuda").scatter_reduce(0, (z * gy + y) * gx + x, dist, "amin")
    assert float(first[first < float("inf")].max()) == 0.0
    # idempotence: fusing one point per occupied voxel (its centre, its orientation) reproduces the volume
    zz, yy, xx = torch.nonzero(occ, as_tuple=True)
    centres = torch.stack([xx, yy, zz], 1).double() * P.VOXEL_SIZE + torch.tensor(P.VOXEL_MIN, device="cuda")
    centres = (centres * torch.tensor([1.0, -1.0, -1.0], device="cuda", dtype=torch.float64)).float()
    world_dirs = vol[..., :3][zz, yy, xx] * torch.tensor([1.0, -1.0, -1.0], device="cuda")
    vol2 = P.voxel_fuse(centres, world_dirs, "cuda:0")
    assert torch.equal(vol2, vol)
