"""CPU-only: the C-ABI library loads and exports every symbol include/monohair_b200.h declares (no compute calls),
and the host-side hooks that need no GPU (torch.topk order emulation) match torch."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from monohair_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "monohair_b200.h")).read()
    declared = set(re.findall(r"\b(mh_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    L = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert set(_lib.exported_symbols()) == declared
    assert _lib.lib().mh_version() >= 100


def test_error_reporting_without_gpu():
    L = _lib.lib()
    rc = L.mh_debug_topk_host(None, 0, 0, None, None)
    assert rc != 0 and b"mh_debug_topk_host" in L.mh_last_error()


@pytest.mark.parametrize("V,k", [(20, 20), (22, 20), (24, 20), (60, 20), (61, 20), (100, 20), (200, 20), (1300, 20), (64, 1), (33, 7)])
def test_topk_order_matches_torch_cpu(V, k):
    """mh_topk.cuh against torch.topk on tie-heavy columns (quantised confidences, many exact zeros)."""
    L = _lib.lib()
    g = torch.Generator().manual_seed(V * 1000 + k)
    N = 400
    lv = torch.randint(0, 256, (V, N), generator=g).float() / 255.0
    few = torch.randint(0, 6, (V, N), generator=g).float() / 5.0
    x = torch.where(torch.rand((V, N), generator=g) < 0.5, torch.zeros(()), lv)
    x[:, N // 2:] = torch.where(torch.rand((V, N - N // 2), generator=g) < 0.3, few[:, N // 2:], x[:, N // 2:])
    x[:, :8] = 0.0                                       # all-equal columns
    x[:, 8:16] = torch.arange(V).float()[:, None]        # sorted ascending / descending
    x[:, 16:24] = -torch.arange(V).float()[:, None]
    val, idx = torch.topk(x, k, dim=0, largest=True)
    out_i = np.empty(k, np.int32)
    out_v = np.empty(k, np.float32)
    for n in range(N):
        col = np.ascontiguousarray(x[:, n].numpy())
        rc = L.mh_debug_topk_host(col.ctypes.data_as(C.c_void_p), V, k, out_i.ctypes.data_as(C.c_void_p),
                                  out_v.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert np.array_equal(out_v, val[:, n].numpy()), f"values differ in column {n}"
        assert np.array_equal(out_i, idx[:, n].numpy().astype(np.int32)), f"tie order differs in column {n}"


def test_sample_offsets_match_oracle():
    from monohair_b200.pmvo import PMVO
    from oracle import pmvo_oracle as O
    assert torch.equal(PMVO._sample_offsets(90), O.sample_offsets(90))
    assert PMVO._sample_offsets(90).numel() == 90


def test_strand_split_matches_per_strand_slicing():
    """HairGrowing._split (device-side compaction + one torch.split) == slicing strand by strand, for the packed layout
    of the segment passes and the fixed-stride layout of the scalp pass; runs on CPU tensors."""
    from monohair_b200.hairgrow import HairGrowing
    g = torch.Generator().manual_seed(3)
    n = 57
    lengths = torch.where(torch.rand(n, generator=g) < 0.3, torch.zeros(n, dtype=torch.int32),
                          torch.randint(5, 40, (n,), generator=g, dtype=torch.int32))
    keep = (torch.rand(n, generator=g) < 0.6) & (lengths > 0)
    # packed: strands back to back
    offsets = torch.cumsum(lengths.long(), 0) - lengths.long()
    pts = torch.rand((int(lengths.sum()), 3), generator=g)
    got = HairGrowing._split(pts, offsets, lengths, keep)
    want = [pts[offsets[i]: offsets[i] + lengths[i]] for i in range(n) if keep[i]]
    assert len(got) == len(want) and all(torch.equal(a, b) for a, b in zip(got, want))
    # fixed stride: strand i starts at i * stride
    stride = 41
    pts2 = torch.rand((n * stride, 3), generator=g)
    off2 = torch.arange(n, dtype=torch.int64) * stride
    got = HairGrowing._split(pts2, off2, lengths, keep, stride=stride)
    want = [pts2[off2[i]: off2[i] + lengths[i]] for i in range(n) if keep[i]]
    assert len(got) == len(want) and all(torch.equal(a, b) for a, b in zip(got, want))
    assert HairGrowing._split(pts, offsets, lengths, torch.zeros(n, dtype=torch.bool)) == []


def test_scalp_normals_interpolate_vertex_normals(tmp_path):
    """HairGrow.py:879-881: open3d returns the barycentric blend of the OBJ's `vn` records (use_triangle_normal=False),
    not the face normal, whose sign would follow the face winding."""
    import numpy as np
    from monohair_b200.pmvo_utils import read_obj_normals, sample_points_uniformly
    p = tmp_path / "tri.obj"
    # one triangle wound so that its face normal is -z, with vn records pointing to +z-ish directions
    p.write_text("v 0 0 0\nv 0 1 0\nv 1 0 0\nvn 0 0 1\nvn 0 0.6 0.8\nvn 0.6 0 0.8\nf 1//1 2//2 3//3\n")
    v, f, vn = read_obj_normals(str(p))
    assert np.allclose(vn, [[0, 0, 1], [0, 0.6, 0.8], [0.6, 0, 0.8]])
    pts, nrm = sample_points_uniformly(v, f, 500, rng=np.random.default_rng(1), with_normals=True, vertex_normals=vn)
    assert (nrm[:, 2] > 0.79).all()                                    # never the (-z) face normal
    # the blend uses the same barycentric weights as the position: for this triangle point = (w2, w1, 0)
    w1, w2 = pts[:, 1], pts[:, 0]
    expect = (1 - w1 - w2)[:, None] * vn[0] + w1[:, None] * vn[1] + w2[:, None] * vn[2]
    assert np.allclose(nrm, expect, atol=1e-12)
    # no vn records: area-weighted vertex normals
    q = tmp_path / "quad.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n")
    v, f, vn = read_obj_normals(str(q))
    assert np.allclose(vn, [[0, 0, 1]] * 4)


def test_connect_strands_and_chain_following_host_logic():
    """HairGrow.py:303-421 host side: connect_strands extends a piece list by the SHAPE of the partner (its successive
    differences) from the free end, through the mid point with the partner's facing end; connect_segments follows the
    connection table through each partner's other end and stops at a strand already on the chain."""
    import numpy as np
    from monohair_b200.hairgrow_connect import connect_segments, connect_strands
    a = np.array([[0., 0, 0], [1, 0, 0], [2, 0, 0]])
    b = np.array([[2.2, 0, 0], [3.2, 1, 0], [4.2, 1, 0]])
    out = connect_strands([a.copy()], b, push_back=True)
    assert len(out) == 2 and out[1].shape == (3, 3)
    mid = a[-1] * 0.5 + b[0] * 0.5
    assert np.allclose(out[1][0], mid) and np.allclose(np.diff(out[1], axis=0), np.diff(b, axis=0))
    out = connect_strands([a.copy()], b[::-1], push_back=False)          # prepend: partner's tip faces our root
    assert len(out) == 2 and np.allclose(out[0][-1], a[0] * 0.5 + b[::-1][-1] * 0.5)
    # three strands in a row: 0.tip -> 1.root, 1.tip -> 2.root; info = {root partner, its end, tip partner, its end}
    s0, s1, s2 = a, a + [2.1, 0, 0], a + [4.2, 0, 0]
    info = np.array([[-1, 0, 1, 1], [0, 2, 2, 1], [1, 2, -1, 0]], np.int32)
    long0 = connect_segments(info, [s0, s1, s2], 0)
    assert long0.shape[0] == 9 and np.all(np.diff(long0[:, 0]) > 0)      # grown through both partners, monotone along x
    long1 = connect_segments(info, [s0, s1, s2], 1)
    assert long1.shape[0] == 9 and np.all(np.diff(long1[:, 0]) > 0)
    # a cycle: the chain check only guards the recursive step (HairGrow.py:331-333), so both of strand 0's own ends still
    # take their partner, but neither side continues back into strand 0
    cyc = np.array([[1, 2, 1, 1], [0, 2, 0, 1]], np.int32)
    assert connect_segments(cyc, [s0, s1], 0).shape[0] == 9


@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (63, 3), (64, 2), (1000, 3), (926011, 8), (5400, 16)])
def test_distributed_sweep_ownership_partitions_the_points(n, world):
    """mh_refine_sweep_dist: rank r owns every world-th block of 64 consecutive points, in increasing order; the host
    index helper and the library's count agree and the ranks' sets partition 0..n-1."""
    from monohair_b200 import pipeline
    from monohair_b200._lib import lib
    L = lib()
    B = int(L.mh_refine_sweep_dist_block())
    seen = []
    for r in range(world):
        idx = pipeline.block_cyclic_index(n, r, world, B, "cpu")
        assert idx.numel() == L.mh_refine_sweep_dist_local_count(n, r, world)
        assert bool((idx[1:] > idx[:-1]).all())
        q = torch.arange(idx.numel())
        assert torch.equal(idx, ((q // B) * world + r) * B + q % B)          # the kernel's index formula
        seen.append(idx)
    assert torch.equal(torch.sort(torch.cat(seen)).values, torch.arange(n))


def test_symmetric_memory_failure_falls_back_once_and_for_all(monkeypatch):
    """pipeline._symm_buffers: when symmetric memory cannot be set up the ranks agree on it through one all-reduce, the
    failure is remembered (no second collective attempt) and the caller gets None -> replicated sweep."""
    import types
    import torch.distributed._symmetric_memory as symm_mem
    from monohair_b200 import pipeline
    calls = []

    class FakeDist:
        ReduceOp = types.SimpleNamespace(MIN="min")
        group = types.SimpleNamespace(WORLD=None)

        def all_reduce(self, t, op=None):
            calls.append(int(t.item()))

    def boom(*a, **k):
        raise RuntimeError("no peer access on this box")
    monkeypatch.setattr(symm_mem, "empty", boom)
    monkeypatch.setattr(pipeline, "_SYMM", {})
    monkeypatch.setattr(pipeline, "_SYMM_BROKEN", False)
    with pytest.warns(UserWarning, match="symmetric memory unavailable"):
        assert pipeline._symm_buffers(100, "cpu", FakeDist()) is None
    assert calls == [0] and pipeline._SYMM_BROKEN
    assert pipeline._symm_buffers(100, "cpu", FakeDist()) is None and calls == [0]


def test_product_never_reaches_for_the_oracle_or_a_cpu_fallback():
    """oracle/ is test infrastructure: nothing the product ships may import it, and the library loader has no fallback."""
    import glob
    shipped = glob.glob(os.path.join(ROOT, "monohair_b200", "**", "*.py"), recursive=True) + \
        [os.path.join(ROOT, f) for f in ("PMVO.py", "HairGrow.py", "options.py")] + \
        glob.glob(os.path.join(ROOT, "Utils", "*.py")) + glob.glob(os.path.join(ROOT, "preprocess_capture_data", "*.py"))
    assert len(shipped) > 10
    for f in shipped:
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle/"
    loader = open(os.path.join(ROOT, "monohair_b200", "_lib.py")).read()
    assert "no CPU fallback" in loader and "raise MonoHairError" in loader
