"""Golden vectors for strand smoothing from the UNMODIFIED reference (Utils/Utils.py smooth_strands, :1148-1198).
    python tests/golden/make_golden_smooth.py      (build container only: needs /root/reference)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402


def strands(seed=0):
    rng = np.random.default_rng(seed)
    lens = [2, 3, 4, 5, 8, 33, 100, 513] + list(rng.integers(5, 120, 40))
    out = []
    for n in lens:
        step = rng.normal(size=(int(n), 3)) * 0.0025                    # ~voxel-sized steps in metres
        out.append((rng.uniform(-0.2, 0.2, 3) + np.cumsum(step, 0)).astype(np.float32))
    return out


def main():
    U = ref_import.import_reference()["Utils"]
    s = strands()
    a = U.smooth_strands([x.copy() for x in s], 4.0, 2.0)                  # HairGrow.py:914 parameters
    b = U.smooth_strands([x.copy() for x in s], 2.0, 1.0, True)            # defaults + fix_tips
    import scipy
    np.savez_compressed(os.path.join(HERE, "smooth_small.npz"), pts=np.concatenate(s, 0),
                        lengths=np.array([x.shape[0] for x in s], np.int32), out_4_2=np.concatenate(a, 0),
                        out_2_1_fix=np.concatenate(b, 0), versions=np.array([np.__version__, scipy.__version__]))
    print("wrote smooth_small.npz:", len(s), "strands,", sum(x.shape[0] for x in s), "points")


if __name__ == "__main__":
    main()
