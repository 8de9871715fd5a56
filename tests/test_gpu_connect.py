"""GPU: HairGrow connect stages against goldens from the UNMODIFIED reference (tests/golden/make_golden_connect.py):
find_connect_info + connect_segments (HairGrow.py:303-546) -> the strands of strands.hair, connect_to_scalp (:606-784) ->
the strands of connected_strands.hair.  numpy's global RNG is seeded as the generator seeded it."""
import numpy as np
import pytest
import torch

from golden_util import load

pytestmark = pytest.mark.gpu
GRID = (256, 256, 192)


@pytest.fixture(scope="module")
def case():
    from monohair_b200.hairgrow import HairGrowing
    g = load("connect_small")
    gx, gy, gz = GRID
    vol = torch.zeros((gz, gy, gx, 4), dtype=torch.float32)
    nz = g["occ_nz"].astype(np.int64)
    o = torch.from_numpy(g["ori_nz"])
    vol[nz[:, 2], nz[:, 1], nz[:, 0], 0] = o[:, 0]
    vol[nz[:, 2], nz[:, 1], nz[:, 0], 1] = -o[:, 1]
    vol[nz[:, 2], nz[:, 1], nz[:, 0], 2] = -o[:, 2]
    vol[nz[:, 2], nz[:, 1], nz[:, 0], 3] = 1.0
    return g, HairGrowing(volume=vol.cuda(), device="cuda:0")


def _split(pts, lens):
    return np.split(pts, np.cumsum(lens)[:-1])


def test_find_connect_info_vs_reference_golden(case, tmp_path):
    from monohair_b200.hairgrow import load_strand, save_hair_strands
    g, hg = case
    bust, num_root = g["bust"], int(g["num_root"])
    # the reference's __main__ flow: voxel strands -> world -> scalp_segment.hair -> load -> + bust for the segments
    world = hg.VoxelToWorld([torch.from_numpy(s.copy()).cuda() for s in _split(g["in_pts"], g["in_len"])], bust)
    save_hair_strands(str(tmp_path / "scalp_segment.hair"), world)
    segment, points = load_strand(str(tmp_path / "scalp_segment.hair"))
    strands, beg = [], 0
    for i, seg in enumerate(segment):
        s = points[beg:beg + seg]
        if i >= num_root:
            s += bust
        strands.append(s)
        beg += seg
    np.random.seed(123)
    connected = hg.find_connect_info(strands[num_root:], float(g["thr"]), float(g["dot_thr"]), hg.occ)
    new_strands = strands[:num_root] + [c - bust for c in connected]
    lens = np.array([s.shape[0] for s in new_strands], np.int32)
    assert np.array_equal(lens, g["a_len"]), f"{int((lens != g['a_len']).sum())} of {len(lens)} connected strands have another length"
    pts = np.concatenate(new_strands, 0)
    same = np.all(pts == g["a_pts"], axis=1)
    print(f"\nconnect_segments: {len(lens)} strands, {int((lens != g['in_len']).sum())} extended; points bit-identical {same.mean() * 100:.3f}%")
    assert same.all()


def test_connect_to_scalp_vs_reference_golden(case):
    g, hg = case
    bust, num_root = g["bust"], int(g["num_root"])
    strands = [s.astype(np.float64) for s in _split(g["b_in_pts"], g["b_in_len"])]      # load_strand gives float64
    # WorldToVoxel (HairGrow.py:826-835) with torch on the CPU, where the golden was made: `x / 0.0025` on a float32 CUDA
    # tensor is evaluated by torch as x * (1 / 0.0025), an ulp away from the CPU's division for some x, and the stage
    # below is discrete (nearest-point indices)
    from monohair_b200.hairgrow import points_to_voxel
    strands = [points_to_voxel(torch.from_numpy(s + bust).type(torch.float)).numpy() for s in strands]
    np.random.seed(321)
    cs = hg.connect_to_scalp(strands, num_root, float(g["out_ratio"]), True)
    lens = np.array([s.shape[0] for s in cs], np.int32)
    assert np.array_equal(lens, g["b_len"]), (len(lens), len(g["b_len"]), int((lens[:min(len(lens), len(g['b_len']))] != g['b_len'][:min(len(lens), len(g['b_len']))]).sum()))
    pts = np.concatenate(cs, 0)
    same = np.all(pts == g["b_pts"], axis=1)
    print(f"\nconnect_to_scalp: {len(lens)} strands kept, points bit-identical {same.mean() * 100:.3f}%")
    assert same.all()
