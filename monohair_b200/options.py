"""YAML configuration surface of the reference (options.py): `--yaml=<path without .yaml>`, `_parent_` inheritance,
`--a.b.c=value` / `--flag` / `--flag!` / `--key=` command-line overrides, seeding, device selection.

Differences: no easydict / termcolor dependency (a small attribute dict is used) and the two interactive prompts of
the reference (options.py:86-93, :116-137) answer 'y' automatically when stdin is not a terminal, so headless runs
do not block.
"""
from __future__ import annotations

import os
import random
import string
import sys

import numpy as np
import torch
import yaml


class edict(dict):
    """attribute-style dict (stands in for easydict.EasyDict, options.py:7)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, edict):
            v = edict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def to_dict(D, dict_type=dict):
    D = dict_type(D)
    for k, v in D.items():
        if isinstance(v, dict):
            D[k] = to_dict(v, dict_type)
        elif isinstance(v, np.ndarray):
            D[k] = v.tolist()
    return D


def parse_arguments(args):
    """options.py:23-46."""
    opt_cmd = {}
    for arg in args:
        assert arg.startswith("--")
        if "=" not in arg[2:]:
            key_str, value = (arg[2:-1], "false") if arg[-1] == "!" else (arg[2:], "true")
        else:
            key_str, value = arg[2:].split("=")
        keys_sub = key_str.split(".")
        opt_sub = opt_cmd
        for k in keys_sub[:-1]:
            if k not in opt_sub:
                opt_sub[k] = {}
            opt_sub = opt_sub[k]
        assert keys_sub[-1] not in opt_sub, keys_sub[-1]
        opt_sub[keys_sub[-1]] = yaml.safe_load(value)
    return edict(opt_cmd)


def _ask(prompt):
    if sys.stdin is None or not sys.stdin.isatty():
        print(prompt + "y  (non-interactive)")
        return "y"
    ans = None
    while ans not in ["y", "n"]:
        ans = input(prompt)
    return ans


def load_options(fname):
    """options.py:62-76."""
    with open(fname) as file:
        opt = edict(yaml.safe_load(file))
    if "_parent_" in opt:
        parent_fnames = opt.pop("_parent_")
        if type(parent_fnames) is str:
            parent_fnames = [parent_fnames]
        for parent_fname in parent_fnames:
            opt_parent = load_options(parent_fname)
            opt_parent = override_options(opt_parent, opt, key_stack=[])
            opt = opt_parent
    print("loading {}...".format(fname))
    return opt


def override_options(opt, opt_over, key_stack=None, safe_check=False):
    """options.py:78-95."""
    for key, value in opt_over.items():
        if isinstance(value, dict):
            opt[key] = override_options(opt.get(key, edict()), value, key_stack=key_stack + [key], safe_check=safe_check)
        else:
            if safe_check and key not in opt:
                key_str = ".".join(key_stack + [key])
                if _ask("\"{}\" not found in original opt, add? (y/n) ".format(key_str)) == "n":
                    print("safe exiting...")
                    sys.exit()
            opt[key] = value
    return opt


def process_options(opt):
    """options.py:97-113."""
    if opt.seed is not None:
        random.seed(opt.seed)
        np.random.seed(opt.seed)
        torch.manual_seed(opt.seed)
        torch.cuda.manual_seed_all(opt.seed)
        if opt.seed != 0:
            opt.name = str(opt.name) + "_seed{}".format(opt.seed)
    else:
        randkey = "".join(random.choice(string.ascii_uppercase) for _ in range(4))
        opt.name = str(opt.name) + "_{}".format(randkey)
    assert isinstance(opt.gpu, int)
    if "LOCAL_RANK" in os.environ:                       # one process per GPU under torchrun
        opt.gpu = int(os.environ["LOCAL_RANK"])
    opt.device = "cpu" if opt.cpu or not torch.cuda.is_available() else "cuda:{}".format(opt.gpu)


def set(opt_cmd={}):
    """options.py:48-60."""
    assert "yaml" in opt_cmd
    fname = "{}.yaml".format(opt_cmd.yaml)
    opt_base = load_options(fname)
    opt = override_options(opt_base, opt_cmd, key_stack=[], safe_check=True)
    process_options(opt)
    return opt


def save_options_file(opt):
    """options.py:116-137."""
    opt_fname = "{}/options.yaml".format(opt.output_path)
    if os.path.isfile(opt_fname):
        with open(opt_fname) as file:
            opt_old = yaml.safe_load(file)
        if to_dict(opt) != opt_old:
            print("existing options file found (different from current one)...")
            if _ask("override? (y/n) ") == "n":
                print("safe exiting...")
                sys.exit()
        else:
            print("existing options file found (identical)")
    else:
        print("(creating new options file...)")
    with open(opt_fname, "w") as file:
        yaml.safe_dump(to_dict(opt), file, default_flow_style=False, indent=4)
