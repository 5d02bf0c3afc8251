"""Minimal stand-in for the `lightning` package (test infrastructure only).

The reference imports `lightning` at module import time (lit_llama/utils.py:15,
generate.py:9) although the decode path never uses it.  `lightning` is not
installed in this image, so golden-vector generation (oracle/make_golden.py)
puts this directory on sys.path.  Nothing in the product imports it.
"""
import random

import numpy as np
import torch

from . import fabric  # noqa: F401


def seed_everything(seed: int) -> int:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


class Fabric:  # only what generate.main() touches
    def __init__(self, devices=1, precision="32-true", **kw):
        self.device = torch.device("cpu")
        self.precision = precision

    def init_module(self, empty_init=False):
        import contextlib

        return contextlib.nullcontext()

    def setup(self, model):
        return model
