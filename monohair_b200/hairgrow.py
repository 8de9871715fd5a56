"""B200-native HairGrow strand generation: host-side mirror of the reference's HairGrow.py (class HairGrowing,
trace / traceFromScalp / GenerateGuideStrandFromScalp / randomlyGenerateSegments, HairGrow.py:40-299).

All seeds are traced in parallel (strand geometry does not depend on the `flag` volume); the
reference's sequential flag gating is reproduced afterwards by an ordered acceptance kernel.  Random jitter can be
injected (`jitter=` [passes, M, 3] uniform [0,1) draws) so runs can be compared draw-for-draw with the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

VOXEL_MIN = (-0.32, -0.32, -0.24)
VOXEL_SIZE = 0.005 / 2
MAX_STEPS = 256            # HairGrow.py:104,142,214
MAX_INNER = 25             # HairGrow.py:216
MIN_LEN = 5                # HairGrow.py:144


def points_to_voxel(points):
    """HairGrow.py:22-28 (flips y,z IN PLACE like the reference)."""
    voxel_min = torch.tensor(VOXEL_MIN, dtype=torch.float, device=points.device)
    points[..., 1:] *= -1
    return (points - voxel_min) / VOXEL_SIZE


def voxel_to_points(voxels):
    """HairGrow.py:30-36."""
    voxel_min = torch.tensor(VOXEL_MIN, dtype=torch.float, device=voxels.device)
    points = voxels * VOXEL_SIZE + voxel_min
    points[..., 1:] *= -1
    return points


class HairGrowing:
    def __init__(self, occ_path=None, ori_path=None, device='cuda:0', image_size=[1120, 1992], volume=None):
        """HairGrow.py:41-55.  Either the two .mat paths (reference signature) or an already fused device
        `volume` float4 [gz,gy,gx,4] straight from the PMVO stage (no 403 MB .mat round trip)."""
        lib()
        self.device = torch.device(device)
        self.image_size = image_size
        if volume is None:
            import scipy.io
            occ = scipy.io.loadmat(occ_path, verify_compressed_data_integrity=False)['Occ']     # [Y,X,Z]
            ori = scipy.io.loadmat(ori_path, verify_compressed_data_integrity=False)['Ori']     # [Y,X,3Z]
            gy, gx, gz = occ.shape
            occ_d = torch.from_numpy(np.ascontiguousarray(occ, dtype=np.float64)).to(self.device)
            ori_d = torch.from_numpy(np.ascontiguousarray(ori, dtype=np.float64)).to(self.device)
            volume = torch.empty((gz, gy, gx, 4), dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                check(lib().mh_volume_from_mat(stream_ptr(self.device), ptr(occ_d), ptr(ori_d), gx, gy, gz, ptr(volume)),
                      "mh_volume_from_mat")
        self.volume = volume.contiguous()
        self.gz, self.gy, self.gx = self.volume.shape[:3]

    # the reference's attributes, as views of the fused volume
    @property
    def occ(self):
        return self.volume[..., 3][None]                      # [1,Z,Y,X]

    @property
    def ori(self):
        return self.volume[..., :3].permute(3, 0, 1, 2)       # [3,Z,Y,X]

    # ------------------------------------------------------------------ batched kernels
    def _trace_batch(self, seeds, thrDot):
        """seeds [n,3] float32 (already jittered).  -> (points [T,3], offsets int64 [n], lengths int32 [n])."""
        n = seeds.size(0)
        dev = self.device
        nf = torch.empty((n,), dtype=torch.int32, device=dev)
        nb = torch.empty((n,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            check(lib().mh_trace_count(st, ptr(self.volume), self.gx, self.gy, self.gz, ptr(seeds), n, float(thrDot),
                                       MAX_STEPS, ptr(nf), ptr(nb)), "mh_trace_count")
            total = nf + nb + 1
            lengths = torch.where(total >= MIN_LEN, total, torch.zeros_like(total))
            offsets = (torch.cumsum(lengths.long(), 0) - lengths.long()).contiguous()
            T = int(lengths.sum().item())
            pts = torch.empty((max(T, 1), 3), dtype=torch.float32, device=dev)
            check(lib().mh_trace_write(st, ptr(self.volume), self.gx, self.gy, self.gz, ptr(seeds), n, float(thrDot),
                                       MAX_STEPS, ptr(nf), ptr(nb), ptr(offsets), MIN_LEN, ptr(pts)), "mh_trace_write")
        return pts, offsets, lengths.contiguous()

    def _accept(self, pts, offsets, lengths, seeds, flag, mode):
        n = lengths.numel()
        acc = torch.empty((n,), dtype=torch.uint8, device=self.device)
        if mode == 0 and seeds is not None and n > 0:
            total = int(pts.size(0))
            wsb = lib().mh_accept_strands_workspace_bytes(n, total)
            ws = torch.empty((wsb,), dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                check(lib().mh_accept_strands_ws(stream_ptr(self.device), ptr(pts), ptr(offsets), ptr(lengths), ptr(seeds), n, total,
                                                 self.gx, self.gy, self.gz, ptr(flag), ptr(acc), ptr(ws), wsb), "mh_accept_strands_ws")
            return acc.bool()
        with torch.cuda.device(self.device):
            check(lib().mh_accept_strands(stream_ptr(self.device), ptr(pts), ptr(offsets), ptr(lengths),
                                          ptr(seeds) if seeds is not None else None, n, self.gx, self.gy, self.gz, mode,
                                          ptr(flag), ptr(acc)), "mh_accept_strands")
        return acc.bool()

    @staticmethod
    def _split(pts, offsets, lengths, keep, stride=None):
        """Kept strands as a list of [L,3] tensors (what the reference's loops build with strands.append): the kept
        points are compacted on the device and cut by ONE torch.split, instead of one Python slice per strand.
        `stride`: points of strand i start at i*stride (scalp batch); None: strands are packed back to back."""
        ln = torch.where(keep, lengths, torch.zeros_like(lengths))
        n = lengths.numel()
        if stride is None:
            total = int(lengths.sum().item())
            sel = torch.repeat_interleave(keep, lengths.long(), output_size=total)
            compact = pts[:total][sel]
        else:
            m = torch.arange(stride, device=pts.device)[None, :] < ln[:, None]
            compact = pts.view(n, stride, 3)[m]
        sizes = ln[keep].cpu().tolist()
        return list(torch.split(compact, sizes)) if sizes else []

    def _scalp_batch(self, roots, normals, thrDot):
        n = roots.size(0)
        dev = self.device
        pts = torch.empty((n, MAX_STEPS + 1, 3), dtype=torch.float32, device=dev)
        ln = torch.empty((n,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib().mh_trace_from_scalp(stream_ptr(dev), ptr(self.volume), self.gx, self.gy, self.gz, ptr(roots),
                                            ptr(normals), n, float(thrDot), MAX_STEPS, MAX_INNER, ptr(pts), ptr(ln)),
                  "mh_trace_from_scalp")
        off = (torch.arange(n, device=dev, dtype=torch.int64) * (MAX_STEPS + 1)).contiguous()
        return pts.view(-1, 3), off, ln

    def _positive_seeds(self):
        """torch.nonzero(occ) in (z,y,x) order flipped to (x,y,z) float (HairGrow.py:230-232)."""
        nz = torch.nonzero(self.volume[..., 3], as_tuple=False)
        return torch.flip(nz, dims=[1]).type(torch.float).contiguous()

    def _segment_passes(self, seeds, flag, thrDot, passes, jitter):
        out = []
        for r in range(passes):
            jit = torch.rand_like(seeds) if jitter is None else torch.as_tensor(jitter[r]).to(self.device).type(torch.float)
            # seedPos += 0.5 ; seedPos += rand*0.5, in place on the seed table
            seeds += torch.tensor([0.5, 0.5, 0.5], dtype=torch.float, device=self.device)
            seeds += jit * 0.5
            pts, off, ln = self._trace_batch(seeds, thrDot)
            keep = self._accept(pts, off, ln, seeds, flag, 0)
            out += self._split(pts, off, ln, keep)
        return out

    # ------------------------------------------------------------------ reference API
    def GenerateGuideStrandFromScalp(self, scalp_points, scalp_normals, pointsTree=None, thrDot=0.8, jitter=None):
        """HairGrow.py:226-265 -> (list of [L,3] device tensors in voxel coordinates, num_root)."""
        print('generate from scalp')
        dev = self.device
        print('voxel size:', self.gz, self.gy, self.gx)
        roots = scalp_points.to(dev).type(torch.float).contiguous()
        normals = scalp_normals.to(dev).type(torch.float).contiguous()
        flag = torch.zeros((self.gz, self.gy, self.gx), dtype=torch.float32, device=dev)
        pts, off, ln = self._scalp_batch(roots, normals, thrDot)
        keep = self._accept(pts, off, ln, None, flag, 1)
        strands = self._split(pts, off, ln, keep, stride=MAX_STEPS + 1)
        print('num guide:', len(strands))
        num_root = len(strands)
        seeds = self._positive_seeds()
        if jitter is not None:
            jitter = np.asarray(jitter).reshape(2, seeds.size(0), 3)
        strands += self._segment_passes(seeds, flag, thrDot, 2, jitter)
        self.strands = strands
        print('done...')
        return strands, num_root

    def randomlyGenerateSegments(self, thrDot=0.8, jitter=None):
        """HairGrow.py:269-299."""
        print('generate segments...')
        flag = torch.zeros((self.gz, self.gy, self.gx), dtype=torch.float32, device=self.device)
        seeds = self._positive_seeds()
        if jitter is not None:
            jitter = np.asarray(jitter).reshape(3, seeds.size(0), 3)
        strands = self._segment_passes(seeds, flag, thrDot, 3, jitter)
        self.strands = strands
        self.strandsTan = []
        print('done...')
        return strands

    def trace(self, seedPos, flag, thrDot, W, H, Z):
        """HairGrow.py:59-149 for one seed: mutates seedPos in place, returns the strand or False."""
        assert (W, H, Z) == (self.gx, self.gy, self.gz)
        seedPos += torch.tensor([0.5, 0.5, 0.5], dtype=torch.float, device=seedPos.device)
        seedPos += torch.rand_like(seedPos) * 0.5
        seeds = seedPos.to(self.device).type(torch.float).reshape(1, 3).contiguous()
        pts, off, ln = self._trace_batch(seeds, thrDot)
        f = flag.to(self.device).type(torch.float).contiguous()
        # the gate only reads `flag`; bumping it stays with the caller as in the reference (:256-260)
        if f[min(max(int(seeds[0, 2]), 0), Z - 1), min(max(int(seeds[0, 1]), 0), H - 1), min(max(int(seeds[0, 0]), 0), W - 1)] >= 3:
            return False
        if int(ln[0]) == 0:
            return False
        return pts[: int(ln[0])]

    def traceFromScalp(self, seedPos, seedNormal, thrDot, W, H, Z, pointsTree=None):
        """HairGrow.py:154-223 for one root: strand or None."""
        assert (W, H, Z) == (self.gx, self.gy, self.gz)
        pts, off, ln = self._scalp_batch(seedPos.to(self.device).type(torch.float).reshape(1, 3).contiguous(),
                                         seedNormal.to(self.device).type(torch.float).reshape(1, 3).contiguous(), thrDot)
        return None if int(ln[0]) == 0 else pts[: int(ln[0])]

    # ------------------------------------------------------------------ connect stages (monohair_b200/hairgrow_connect.py)
    def find_connect_info(self, strands, connect_threshold=0.005, connect_dot_threshold=0.7, occ=None):
        """HairGrow.py:436-546 (the `occ` argument of the reference is this solver's own volume)."""
        from . import hairgrow_connect as HC
        return HC.find_connect_info(strands, connect_threshold, connect_dot_threshold, self.volume, self.device)

    def connect_segments(self, strands_connect_info, strands, i):
        """HairGrow.py:303-346 with the connection table of mh_connect_find (int32 [n,4])."""
        from . import hairgrow_connect as HC
        return HC.connect_segments(strands_connect_info, strands, i)

    def connect_strands(self, strand1, strand2, push_back, cubic_sample=False, add_mid=True, need_weight=False):
        from . import hairgrow_connect as HC
        return HC.connect_strands(strand1, strand2, push_back, cubic_sample, add_mid, need_weight)

    def connect_to_scalp(self, strands, num_root, out_ratio=0.5, infer_inner=True):
        """HairGrow.py:606-784 (out_ratio = args.HairGenerate.out_ratio, a module global in the reference)."""
        from . import hairgrow_connect as HC
        return HC.connect_to_scalp(strands, num_root, self.volume, out_ratio, infer_inner)

    def VoxelToWorld(self, strands, bust_to_origin=None):
        """HairGrow.py:816-824."""
        if len(strands) == 0:
            return []
        if torch.is_tensor(strands[0]):
            # all strands through ONE transform and ONE device->host copy (the reference converts strand by strand)
            sizes = [int(s.shape[0]) for s in strands]
            allp = voxel_to_points(torch.cat(strands, 0)).cpu().numpy()
            if bust_to_origin is not None:
                allp -= np.asarray(bust_to_origin)          # float64 operand, rounded into the float32 array like the reference's ss -= ...
            return np.split(allp, np.cumsum(sizes)[:-1])
        out = []
        for ss in strands:
            ss = voxel_to_points(torch.as_tensor(ss)).cpu().numpy()
            if bust_to_origin is not None:
                ss -= bust_to_origin
            out.append(ss)
        return out

    def WorldToVoxel(self, strands, bust_to_origin=None):
        """HairGrow.py:826-835."""
        out = []
        for ss in strands:
            if bust_to_origin is not None:
                ss += bust_to_origin
            t = torch.from_numpy(ss).type(torch.float).to(self.device)
            out.append(points_to_voxel(t).cpu().numpy())
        return out


def smooth_strands(strands, lap_constraint=2.0, pos_constraint=1.0, fix_tips=False, device="cuda:0"):
    """Utils/Utils.py:1194-1198 (smnooth_strand :1148-1192 per strand) on the device: list of [L,3] arrays (or tensors)
    -> list of smoothed [L,3] float32 numpy arrays (the reference stores the float64 solution back into the strand's
    float32 array).  With fix_tips the end points keep their positions (:1186-1187)."""
    if len(strands) == 0:
        return strands
    dev = torch.device(device)
    arrs = [np.asarray(s.detach().cpu() if torch.is_tensor(s) else s, dtype=np.float32).reshape(-1, 3) for s in strands]
    lengths = np.array([a.shape[0] for a in arrs], dtype=np.int32)
    offsets = (np.cumsum(lengths.astype(np.int64)) - lengths).astype(np.int64)
    total = int(lengths.sum())
    pts = torch.from_numpy(np.ascontiguousarray(np.concatenate(arrs, 0))).to(dev)
    out = torch.empty_like(pts)
    wsb = lib().mh_smooth_strands_workspace_bytes(total)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    t_off, t_len = torch.from_numpy(offsets).to(dev), torch.from_numpy(lengths).to(dev)
    with torch.cuda.device(dev):
        check(lib().mh_smooth_strands(stream_ptr(dev), ptr(pts), ptr(t_off), ptr(t_len), len(arrs), float(lap_constraint),
                                      float(pos_constraint), ptr(out), ptr(ws), wsb, total), "mh_smooth_strands")
    res = out.cpu().numpy()
    smoothed = []
    for a, o, n in zip(arrs, offsets, lengths):
        sm = res[o:o + n].copy()
        if fix_tips and n >= 2:
            tips = a.copy()
            tips[1:-1] = sm[1:-1]
            sm = tips
        smoothed.append(sm)
    return smoothed


def save_hair_strands(path, strands):
    """Utils/Utils.py:1246-1262: uint32 n_strands, uint32 n_points, uint16[n_strands], float32[n_points*3]."""
    segments = np.array([s.shape[0] for s in strands], dtype=np.uint16)
    pts = np.concatenate(strands, 0).astype(np.float32) if len(strands) else np.zeros((0, 3), np.float32)
    with open(path, 'wb') as f:
        f.write(np.array([len(strands), pts.shape[0]], dtype=np.uint32).tobytes())
        f.write(segments.tobytes())
        f.write(np.ascontiguousarray(pts).tobytes())


def load_strand(file):
    """Utils/PMVO_utils.py:47-66 -> (segments list, points [n,3] float64)."""
    with open(file, 'rb') as f:
        n_strand, n_pts = np.frombuffer(f.read(8), dtype=np.uint32)
        segments = np.frombuffer(f.read(2 * int(n_strand)), dtype=np.uint16)
        pts = np.frombuffer(f.read(4 * int(segments.sum()) * 3), dtype=np.float32)
    return list(int(s) for s in segments), pts.astype(np.float64).reshape(-1, 3)
