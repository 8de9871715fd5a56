"""HairGrow connect stages: host side (mirror of HairGrowing.find_connect_info / connect_segments / connect_strands /
connect_to_scalp, /root/reference HairGrow.py:303-546, :606-811, and Utils/PMVO_utils.py:random_move_strands :618-658).

Stage A (`find_connect_info`, -> strands.hair): the search for each strand's connection partners -- the part that is
KDTree queries in a Python loop over 1e5 strands in the reference -- runs as one CUDA kernel (csrc/connect.cu,
mh_connect_find); the first occupancy test of every connected strand is one batched kernel (mh_strand_occupancy).  What
stays on the host is what is sequential by construction in the reference: following the connection chains, and the
retry loop of the few strands that fail the occupancy test, because it consumes numpy's global RNG in strand order.

Stage B (`connect_to_scalp`, -> connected_strands.hair) is an order-dependent host algorithm in the reference (its result
depends on the traversal order of scipy's KDTree.query_ball_point and on strands being rewritten while the loop runs); it
is kept a host algorithm here, over the same scipy trees, with the volume look-ups on a host copy of the fused volume.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import MonoHairError, check, lib, ptr, stream_ptr

_OTHER_END = {1: 2, 2: 1}        # 1 = root, 2 = tip (mh_connect_find)


def _pack(strands, dev):
    lengths = np.array([s.shape[0] for s in strands], dtype=np.int32)
    offsets = (np.cumsum(lengths.astype(np.int64)) - lengths).astype(np.int64)
    pts = np.ascontiguousarray(np.concatenate(strands, 0), dtype=np.float64) if len(strands) else np.zeros((0, 3))
    return (torch.from_numpy(pts).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(lengths).to(dev))


def connect_info(strands, connect_threshold, connect_dot_threshold, device):
    """mh_connect_find -> int32 [n,4] on the host: {root partner, its end, tip partner, its end} (-1 / 0 = none)."""
    dev = torch.device(device)
    n = len(strands)
    if n == 0:
        return np.zeros((0, 4), np.int32)
    pts, off, ln = _pack(strands, dev)
    info = torch.empty((n, 4), dtype=torch.int32, device=dev)
    ovf = torch.zeros((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().mh_connect_find(stream_ptr(dev), ptr(pts), ptr(off), ptr(ln), n, float(connect_threshold),
                                    float(connect_dot_threshold), ptr(info), ptr(ovf)), "mh_connect_find")
    if int(ovf.item()):
        raise MonoHairError("mh_connect_find: more than 256 strand ends inside one connect_threshold radius")
    return info.cpu().numpy()


def occupancy_fraction(strands, volume, device, shift=None, voxel_space=False):
    """mh_strand_occupancy -> float64 [n] on the host (share of points on occupied voxels, -1 = left the grid)."""
    dev = torch.device(device)
    n = len(strands)
    if n == 0:
        return np.zeros((0,))
    pts, off, ln = _pack(strands, dev)
    frac = torch.empty((n,), dtype=torch.float64, device=dev)
    sh = None if shift is None else torch.from_numpy(np.ascontiguousarray(shift, dtype=np.float64)).to(dev)
    gz, gy, gx = volume.shape[:3]
    with torch.cuda.device(dev):
        check(lib().mh_strand_occupancy(stream_ptr(dev), ptr(pts), ptr(off), ptr(ln), n, ptr(sh), ptr(volume), gx, gy, gz,
                                        1 if voxel_space else 0, ptr(frac)), "mh_strand_occupancy")
    return frac.cpu().numpy()


def connect_strands(strand1, strand2, push_back, cubic_sample=False, add_mid=True, need_weight=False):
    """HairGrow.py:349-421: extend the piece list `strand1` by the SHAPE of strand2 (its successive differences), starting at
    the free end of strand1 (optionally through the mid point with strand2's facing end).  Returns the same list."""
    n = strand2.shape[0]
    if push_back:
        seed = strand1[-1][-1]
        piece = []
        if add_mid:
            seed = seed * 0.5 + strand2[0] * 0.5
            piece.append(seed[None])
        for i in range(n - 1):
            nxt = seed + (strand2[i + 1] - strand2[i])
            nxt = nxt * (1 - 0) + strand2[i + 1] * 0          # weight = 0 (:372-373)
            piece.append(nxt[None])
            seed = nxt
        strand1.append(np.concatenate(piece, 0))
    else:
        seed = strand1[0][0]
        piece = []
        if add_mid:
            seed = seed * 0.5 + strand2[-1] * 0.5
            piece.append(seed[None])
        for i in range(n - 1):
            nxt = seed + (strand2[-2 - i] - strand2[-1 - i])
            w = np.sin(0.5 * np.pi * min(1, i / (1.5 * n))) if need_weight else 0
            nxt = nxt * (1 - w) + strand2[-2 - i] * w
            piece.append(nxt[None])
            seed = nxt
        strand1.insert(0, np.concatenate(piece, 0)[::-1])
    return strand1


def connect_segments(info, strands, i):
    """HairGrow.py:303-346: strand i grown through its chain of partners, root side first (each partner is entered through
    the end the connection points at and left through its other end; a strand already on the chain stops it)."""
    pieces = [strands[i]]
    chain = [i]

    def follow(best, end, along_with_root):
        nonlocal pieces
        while True:
            chain.append(best)
            s = strands[best]
            if end == 1:        # the partner's root faces us
                pieces = connect_strands(pieces, s[::-1], False) if along_with_root else connect_strands(pieces, s, True)
            else:               # its tip faces us
                pieces = connect_strands(pieces, s, False) if along_with_root else connect_strands(pieces, s[::-1], True)
            other = _OTHER_END[end]
            nb, ne = (info[best][0], info[best][1]) if other == 1 else (info[best][2], info[best][3])
            if nb < 0 or nb in chain:
                return
            best, end = int(nb), int(ne)

    if info[i][0] >= 0:
        follow(int(info[i][0]), int(info[i][1]), True)
    if info[i][2] >= 0:
        follow(int(info[i][2]), int(info[i][3]), False)
    return np.concatenate(pieces, 0)


def find_connect_info(strands, connect_threshold=0.005, connect_dot_threshold=0.7, volume=None, device="cuda:0"):
    """HairGrow.py:436-546.  strands: list of float64 [L,3] world-frame arrays; volume: the fused float4 volume.
    -> list of connected strands.  numpy's global RNG is consumed exactly as the reference consumes it."""
    print('connect segments...')
    info = connect_info(strands, connect_threshold, connect_dot_threshold, device)
    longs = [connect_segments(info, strands, i) for i in range(len(strands))]
    frac = occupancy_fraction(longs, volume, device)               # first test of every strand, no random numbers involved
    out, fail = [], 0
    for i, strand in enumerate(longs):
        if frac[i] < 0:                                            # left the grid: kept as it is (:517-519)
            fail += 1
        elif not frac[i] > 0.8:
            ok = False
            for count in range(1, 51):                             # :529-536: a fresh random shift of the ORIGINAL strand per retry
                shift = np.random.random((3)) * 0.005
                if count >= 50:
                    break
                f = occupancy_fraction([strand], volume, device, shift=shift[None])[0]
                if f < 0:
                    break
                if f > 0.8:
                    strand = strand + shift
                    ok = True
                    break
            fail += 0 if ok else 1
        out.append(strand)
    print('fail:', fail)
    print('done...')
    return out


# ------------------------------------------------------------------------------------------------------------ stage B
def compute_similar(A, B):
    """Utils/Utils.py:1200-1201."""
    return np.sum(A * B, axis=-1) / (np.maximum(np.linalg.norm(A, 2, axis=-1) * np.linalg.norm(B, 2, axis=-1), 1e-4))


def compute_strands_similar(strand1, strand2, Tan, thr_dist, thr_dot, ori):
    """HairGrow.py:787-811."""
    min_loss, min_index = np.inf, None
    for i in range(strand2.shape[0]):
        dist = np.linalg.norm(strand2[i] - strand1)
        similar_connect = compute_similar(strand1 - strand2[i], Tan)
        similar = compute_similar(ori, Tan)
        if similar > thr_dot and dist < thr_dist + strand2.shape[0] - i - 1:
            loss = (1 - similar_connect) + 0.1 * thr_dist
            if loss < min_loss:
                min_loss, min_index = loss, i
    if min_index is not None:
        min_index = min_index - strand2.shape[0] + 1
    return min_loss, min_index


class _HostVolume:
    """host copy of the fused volume for the strand-by-strand look-ups of stage B (occ [Z,Y,X], ori [3,Z,Y,X] float32 in
    HairGrowing's frame)."""

    def __init__(self, volume):
        v = volume.cpu().numpy()
        self.occ = np.ascontiguousarray(v[..., 3])
        self.ori = np.ascontiguousarray(np.moveaxis(v[..., :3], -1, 0))


def random_move_strands(original_strand, hv, threshold=0.4, index=-1):
    """Utils/PMVO_utils.py:618-658 (its retry count is 1, so no random number is ever drawn)."""
    ss = torch.from_numpy(original_strand.copy()[:index])
    strand_ori = torch.cat([ss[1:] - ss[:-1], ss[-1:] - ss[-2:-1]], 0)
    idx = torch.round(ss).type(torch.long)
    if torch.max(idx[:, 2]) >= 192 or torch.max(idx[:, 1] >= 256) or torch.max(idx[:, 0] >= 256):
        return original_strand, False, 0
    z, y, x = idx[:, 2].numpy(), idx[:, 1].numpy(), idx[:, 0].numpy()
    ss_occ = torch.from_numpy(hv.occ[z, y, x])
    ss_ori = torch.from_numpy(hv.ori[:, z, y, x])
    so = strand_ori.type(ss_ori.dtype) if strand_ori.dtype != ss_ori.dtype else strand_ori
    s1 = torch.cosine_similarity(ss_ori.permute(1, 0), so, dim=-1)
    s2 = torch.cosine_similarity(-ss_ori.permute(1, 0), so, dim=-1)
    similar = torch.sum(torch.maximum(s2, s1)) / torch.sum(ss_occ)
    out_ratio = 1 - (torch.sum(ss_occ) / ss_occ.size(0))
    if torch.sum(ss_occ) / ss_occ.size(0) > threshold and similar > 0.3:
        return original_strand.copy(), True, out_ratio
    return original_strand, False, out_ratio


def connect_to_scalp(strands, num_root, volume, out_ratio_threshold, infer_inner=True):
    """HairGrow.py:606-784.  strands: list of [L,3] arrays in VOXEL coordinates (WorldToVoxel output), the first num_root of
    them rooted on the scalp.  Poor strands are attached to good ones in rounds of growing search radius / shrinking
    orientation threshold.  -> list of strands that ended up rooted (or flagged 'out'), in order."""
    from scipy.spatial import KDTree
    hv = _HostVolume(volume)
    n = len(strands)
    root_flag = np.zeros((n,))
    root_flag[:num_root] = 1
    out_ratio = np.zeros_like(root_flag)
    print('num of strands:', n)
    print('num of good strands:', np.sum(root_flag))
    root_flag = root_flag.astype(np.bool_)
    out_root_flag = np.zeros_like(root_flag).astype(np.bool_)
    print('connect poor strands to good strands...')
    thr_dist, thr_dot, max_thr_dist, max_dot_dist = 0.5, 0.9, 2.0, 0.6       # the infer_inner branch sets the same values (:631-635)
    it, flag = 0, True
    while flag:
        print('iter:', it)
        print('num of good strands:', np.sum(root_flag))
        print('num of out strands:', np.sum(out_root_flag))
        print('current thr_dist:', thr_dist)
        print('current thr_dot:', thr_dot)
        num_good = np.sum(root_flag)
        strands_info, core, trees = [], [], []
        for i in range(n):
            if root_flag[i]:
                core.append(strands[i])
                strands_info.extend([i] * strands[i].shape[0])
            trees.append(KDTree(data=strands[i]))
        strands_info = np.array(strands_info)
        core_tree = KDTree(data=np.concatenate(core, 0))
        for i in range(n):
            if root_flag[i] or out_root_flag[i]:
                continue
            strand = strands[i]
            nei_index = core_tree.query_ball_point(strand[0], thr_dist)
            nei_strands = strands_info[nei_index]
            nei_index_inv = core_tree.query_ball_point(strand[-1], thr_dist * 2)
            nei_strands_inv = strands_info[nei_index_inv]
            if len(nei_strands_inv) != 0 and len(np.union1d(nei_index_inv, nei_index)) == 0:
                continue
            if len(nei_index) != 0:
                closest = nei_strands[0]
                nei_pos_dist, nei_pos_index = trees[closest].query(strand, 1)
                beg, end = nei_pos_index[0], nei_pos_index[-1]
                ss = strands[closest]
                tan1 = ss[beg] - ss[beg - 1] if beg == ss.shape[0] - 1 else ss[beg + 1] - ss[beg]
                tan2 = strand[1] - strand[0]
                if compute_similar(tan1, tan2) < 0 and beg > end and np.mean(nei_pos_dist) < 5:
                    strands[i] = strand[::-1]
                    strand = strands[i]
            connect, min_loss, best_pos_index, min_nei_point_index, min_nei_strandI = False, np.inf, None, None, None
            seen, count = [], 0
            for neiI in nei_strands:
                if neiI in seen:
                    continue
                seen.append(neiI)
                count += 1
                nei_strand = strands[neiI]
                _, nei_point_index = trees[neiI].query(strand[0], 1)
                nei_distance, _ = trees[neiI].query(strand[:5], 1)
                if np.mean(nei_distance) < 1:
                    continue
                if len(strand) > 60 and len(strand) + nei_point_index > 150:
                    continue
                Tan = strand[1] - strand[0]
                if nei_point_index <= 1:
                    continue
                nei_ori = nei_strand[nei_point_index] - nei_strand[nei_point_index - 1]
                loss, pos_index = compute_strands_similar(strand[0], nei_strand[nei_point_index:nei_point_index + 1], Tan, thr_dist,
                                                          thr_dot, nei_ori)
                loss += out_ratio[neiI]
                if loss < min_loss:
                    min_loss, min_nei_strandI, best_pos_index = loss, neiI, pos_index
                    min_nei_point_index = nei_point_index + best_pos_index
                    connect = True
                if count >= 30:
                    break
            if not connect:
                continue
            if min_nei_point_index <= 1 or best_pos_index is None:
                continue
            ss = strands[min_nei_strandI]
            mid_point = strand[0] * 0.95 + ss[min_nei_point_index] * 0.05
            pieces = connect_strands([mid_point[None], strand[0:]], ss[:min_nei_point_index + 1], push_back=False,
                                     cubic_sample=False, add_mid=False)
            connect_strand = np.concatenate(pieces, 0)
            connect_strand, in_check, out_r = random_move_strands(connect_strand, hv, out_ratio_threshold,
                                                                  index=min_nei_point_index + 1)
            out_ratio[i] = out_r
            strands[i] = connect_strand
            if in_check:
                root_flag[i] = True
            else:
                out_root_flag[i] = True
        if np.sum(root_flag) - num_good > (n - num_root) // 500:
            flag = True
        elif thr_dist == max_thr_dist and thr_dot == max_dot_dist:
            flag = False
        else:
            thr_dist = min(thr_dist + 0.25, max_thr_dist)
            thr_dot = max(thr_dot - 0.075, max_dot_dist)
            flag = True
        it += 1
    print('done...')
    print('connect to scalp...')
    return [strands[i] for i in range(n) if root_flag[i] or out_root_flag[i]]
