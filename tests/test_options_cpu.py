"""CPU: YAML surface of the reference (options.py semantics) and the hot-path config keys."""
import os

from monohair_b200 import options

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parent_inheritance_and_cli_overrides(monkeypatch):
    monkeypatch.chdir(ROOT)
    cmd = options.parse_arguments(["--yaml=configs/reconstruct/big_wavy1", "--PMVO.infer_inner", "--PMVO.optimize=",
                                   "--PMVO.threshold=0.03", "--HairGenerate.connect_scalp!"])
    assert cmd.PMVO.infer_inner is True and cmd.PMVO.optimize is None and cmd.HairGenerate.connect_scalp is False
    opt = options.set(opt_cmd=cmd)
    assert opt.PMVO.patch_size == 7 and opt.PMVO.conf_threshold == 0.15          # big_wavy1 overrides
    assert opt.PMVO.visible_threshold == 1 and opt.PMVO.filter_point is True    # inherited from base
    assert opt.PMVO.threshold == 0.03 and opt.PMVO.optimize is None
    assert opt.data.image_size == [1920, 1080] and opt.data.case == "big_wavy1"
    assert opt.HairGenerate.grow_threshold == 0.85
    assert opt.name == "10-16" and opt.device in ("cpu", "cuda:0")
    assert list(opt.bbox_min) == [-0.32, -0.32, -0.24]


def test_save_options_roundtrip(tmp_path, monkeypatch):
    monkeypatch.chdir(ROOT)
    opt = options.set(opt_cmd=options.parse_arguments(["--yaml=configs/reconstruct/big_wavy1"]))
    opt.output_path = str(tmp_path)
    options.save_options_file(opt)
    options.save_options_file(opt)           # identical -> no prompt
    opt.PMVO.threshold = 0.5
    options.save_options_file(opt)           # differs -> auto 'y' when non-interactive
    import yaml
    assert yaml.safe_load(open(tmp_path / "options.yaml"))["PMVO"]["threshold"] == 0.5
