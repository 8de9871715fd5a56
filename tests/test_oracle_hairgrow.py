"""CPU: HairGrow oracle (oracle/hairgrow_oracle.py) against goldens from the unmodified reference (bit-exact)."""
import numpy as np

from golden_util import load
from oracle import hairgrow_oracle as H


def _split(pts, lens):
    out, b = [], 0
    for n in lens:
        out.append(pts[b:b + n])
        b += n
    return out


def test_guide_strands_bit_exact():
    g = load("hairgrow_small")
    vol = H.Volume.from_memory(g["occ"].astype(np.float64), g["ori"].astype(np.float64))
    strands, num_root = H.generate_guide_strands(vol, g["roots"], g["normals"], float(g["thr"]), g["jitter"], passes=2)
    assert num_root == int(g["guide_num_root"])
    assert np.array_equal(np.array([s.shape[0] for s in strands], np.int32), g["guide_len"])
    assert np.array_equal(np.concatenate(strands), g["guide_pts"])


def test_random_segments_bit_exact():
    g = load("hairgrow_small")
    vol = H.Volume.from_memory(g["occ"].astype(np.float64), g["ori"].astype(np.float64))
    strands = H.randomly_generate_segments(vol, float(g["thr"]), g["jitter"], passes=3)
    assert np.array_equal(np.array([s.shape[0] for s in strands], np.int32), g["segments_len"])
    assert np.array_equal(np.concatenate(strands), g["segments_pts"])


def test_smooth_strands_oracle_vs_reference_golden():
    """oracle smooth_strand (scipy sparse LU, as the reference) == the unmodified reference bit for bit; the banded-Cholesky
    formulation the CUDA kernel uses agrees with it on the stored float32 values."""
    g = load("smooth_small")
    ins = _split(g["pts"], g["lengths"])
    for key, lap, pos, fix in (("out_4_2", 4.0, 2.0, False), ("out_2_1_fix", 2.0, 1.0, True)):
        want = _split(g[key], g["lengths"])
        for s, w in zip(ins, want):
            got = H.smooth_strand(s.copy(), lap, pos, fix)
            assert got.dtype == np.float32 and np.array_equal(got, w)
            if not fix:
                banded = H.smooth_strand_banded(s.copy(), lap, pos)
                ulp = np.spacing(np.abs(w).astype(np.float32))
                assert np.all(np.abs(banded - w) <= ulp)
                assert np.mean(banded == w) > 0.999
