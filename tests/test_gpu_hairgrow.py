"""GPU parity of the HairGrow trace kernels against goldens from the unmodified reference and the CPU oracle.
Strand geometry is float32 adds of gathered voxel values in the reference's order: bit-exact is the bar."""
import numpy as np
import pytest
import scipy.io
import torch

from golden_util import load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solver(tmp_path_factory):
    from monohair_b200.hairgrow import HairGrowing
    from oracle import pmvo_oracle as O
    g = load("hairgrow_small")
    td = tmp_path_factory.mktemp("vol")
    mo, mori = O.mat_layout(g["occ"].astype(np.float64), g["ori"].astype(np.float64))
    scipy.io.savemat(str(td / "Occ3D.mat"), {"Occ": mo})
    scipy.io.savemat(str(td / "Ori3D.mat"), {"Ori": mori})
    return g, HairGrowing(str(td / "Occ3D.mat"), str(td / "Ori3D.mat"), device="cuda:0")


def test_volume_reader_layout(solver):
    g, hg = solver
    from oracle import hairgrow_oracle as H
    vol = H.Volume.from_memory(g["occ"].astype(np.float64), g["ori"].astype(np.float64))
    assert np.array_equal(hg.occ[0].cpu().numpy(), vol.occ)
    assert np.array_equal(hg.ori.cpu().numpy(), vol.ori)
    assert (hg.gx, hg.gy, hg.gz) == tuple(int(x) for x in g["grid"])


def test_guide_strands_vs_reference_golden(solver):
    g, hg = solver
    strands, num_root = hg.GenerateGuideStrandFromScalp(torch.from_numpy(g["roots"]), torch.from_numpy(g["normals"]), None,
                                                        float(g["thr"]), jitter=g["jitter"][: 2 * int((g["occ"] > 0).sum())])
    assert num_root == int(g["guide_num_root"])
    lens = np.array([s.shape[0] for s in strands], np.int32)
    assert np.array_equal(lens, g["guide_len"])
    assert np.array_equal(torch.cat(strands).cpu().numpy(), g["guide_pts"])


def test_random_segments_vs_reference_golden(solver):
    g, hg = solver
    strands = hg.randomlyGenerateSegments(float(g["thr"]), jitter=g["jitter"])
    lens = np.array([s.shape[0] for s in strands], np.int32)
    assert np.array_equal(lens, g["segments_len"])
    assert np.array_equal(torch.cat(strands).cpu().numpy(), g["segments_pts"])


def test_single_seed_api_and_hair_codec(solver, tmp_path):
    from monohair_b200.hairgrow import load_strand, save_hair_strands
    g, hg = solver
    seed = torch.tensor([24.0, 14.0, 20.0], device="cuda:0")
    flag = torch.zeros((hg.gz, hg.gy, hg.gx), device="cuda:0")
    before = seed.clone()
    s = hg.trace(seed, flag, 0.85, hg.gx, hg.gy, hg.gz)
    assert torch.all(seed >= before + 0.5) and torch.all(seed < before + 1.0)       # in-place jitter (§9-R8)
    assert s is False or s.shape[1] == 3
    r = hg.traceFromScalp(torch.from_numpy(g["roots"][0]), torch.from_numpy(g["normals"][0]), 0.85, hg.gx, hg.gy, hg.gz)
    assert r is None or np.array_equal(r.cpu().numpy()[0], g["roots"][0])
    strands = [np.random.rand(7, 3).astype(np.float32), np.random.rand(5, 3).astype(np.float32)]
    save_hair_strands(str(tmp_path / "a.hair"), strands)
    seg, pts = load_strand(str(tmp_path / "a.hair"))
    assert seg == [7, 5] and np.array_equal(pts.astype(np.float32), np.concatenate(strands))


def test_full_size_properties():
    """256x256x192 analytic volume: every strand step is a unit-length move between occupied voxels, strands have
    5..513 points, and flag gating keeps at most 3 accepted strands through any seed voxel's first visit."""
    from monohair_b200 import synthetic as syn
    from monohair_b200.hairgrow import HairGrowing
    occ, ori = syn.orientation_volume(device="cuda:0")
    vol = torch.zeros((192, 256, 256, 4), device="cuda:0")
    o = torch.from_numpy(ori).cuda().float()
    vol[..., 0] = o[..., 0].permute(2, 1, 0)
    vol[..., 1] = -o[..., 1].permute(2, 1, 0)
    vol[..., 2] = -o[..., 2].permute(2, 1, 0)
    vol[..., 3] = torch.from_numpy(occ).cuda().float().permute(2, 1, 0)
    hg = HairGrowing(volume=vol, device="cuda:0")
    seeds = hg._positive_seeds()[::50].contiguous() + 0.6
    pts, off, ln = hg._trace_batch(seeds, 0.85)
    ln_c = ln.cpu().numpy()
    assert ((ln_c == 0) | ((ln_c >= 5) & (ln_c <= 513))).all()
    assert (ln_c > 0).mean() > 0.5
    i = int(np.argmax(ln_c))
    s = pts[int(off[i]):int(off[i]) + int(ln[i])]
    step = (s[1:] - s[:-1]).norm(dim=1)
    assert torch.allclose(step, torch.ones_like(step), atol=1e-4)
