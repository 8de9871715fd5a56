"""Load committed golden fixtures (tests/golden/*.npz) back into Scene objects."""
import os

import numpy as np

from monohair_b200.synthetic import Scene

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def scene_of(g):
    cams = [{"file": "view_%03d" % i, "pose": g["poses"][i].tolist(), "ndc_prj": g["ndc_prj"][i].tolist()}
            for i in range(g["poses"].shape[0])]
    return Scene(H=int(g["H"]), W=int(g["W"]), cams=cams, depth=g["depth"], ori_gray=g["ori_gray"],
                 conf_u8=g["conf_u8"], mask_u8=g["mask_u8"])
