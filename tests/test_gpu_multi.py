"""GPU, >= 2 devices (skipped on a 1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
the sharded job equals the single-GPU job bit for bit, in both fusion modes, and the torchrun CLI writes the same files
as the single-process CLI."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import scipy.io
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(script_args, world, env=None, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_port())] + script_args
    return subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT, **(env or {})), capture_output=True, text=True,
                          timeout=timeout)


@needs2
@pytest.mark.parametrize("fusion,sweep", [("replicated", "peer"), ("winners", "replicated")])
def test_sharded_job_equals_single_gpu_job(fusion, sweep):
    world = min(torch.cuda.device_count(), 8)
    r = _torchrun([os.path.join(ROOT, "tests", "multi_gpu_check.py")], world, env={"MH_FUSE_DIST": fusion, "MH_SWEEP_DIST": sweep})
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@needs2
def test_torchrun_cli_writes_the_single_process_files(tmp_path):
    from dataset_util import write_capture
    from monohair_b200 import synthetic as syn
    sc = syn.make_scene(V=24, H=180, W=240, seed=6)
    write_capture(str(tmp_path / "data"), sc, case="synth")
    cfgdir = tmp_path / "cfg"
    cfgdir.mkdir()
    for name in ("one", "two"):
        (cfgdir / f"{name}.yaml").write_text(f"""_parent_: {ROOT}/configs/reconstruct/base.yaml
name: {name}
data:
  root: {tmp_path}/data
  case: synth
  image_size: [180, 240]
PMVO:
  num_sample_per_grid: 2
  patch_size: 5
  threshold: 0.05
  conf_threshold: 0.15
  infer_inner:
""")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "PMVO.py"), f"--yaml={cfgdir}/one"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    r = _torchrun([os.path.join(ROOT, "PMVO.py"), f"--yaml={cfgdir}/two"], 2)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    a, b = tmp_path / "data" / "synth" / "output" / "one", tmp_path / "data" / "synth" / "output" / "two"
    for f in ("optimize/surface.npy", "optimize/filter_unvisible.npy", "optimize/select_p.npy", "optimize/select_o.npy",
              "optimize/min_loss.npy", "optimize/high_conf_index.npy", "refine/select_o.npy", "refine/min_loss.npy",
              "refine/filter_unvisible.npy", "refine/filter_unvisible_ori.npy"):
        x, y = np.load(a / f), np.load(b / f)
        assert x.shape == y.shape and np.array_equal(x, y), f
    for f, k in (("refine/Occ3D.mat", "Occ"), ("refine/Ori3D.mat", "Ori")):
        assert np.array_equal(scipy.io.loadmat(a / f)[k], scipy.io.loadmat(b / f)[k]), f
