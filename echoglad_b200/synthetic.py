"""Synthetic inputs in the shape of the reference's DummyDataset (src/core/datasets.py:1381-1439):
randn frames, 4 integer (h, w) landmarks per frame uniform in [0, S-1], valid = 1.  Labels are produced
on the device from the coordinates (eg_node_labels), replacing create_node_labels (:1586-1612)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .graph import HierGraphSpec


def host_batch(batch: int, frame_size: int, seed: int = 200, pin: bool = True):
    """(frames float32[B,1,S,S], coords int32[B,4,2]) in (pinned) host memory."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.randn(batch, 1, frame_size, frame_size, generator=g)
    rng = np.random.default_rng(seed)
    coords = torch.from_numpy(rng.integers(0, frame_size, size=(batch, 4, 2)).astype(np.int32))
    if pin and torch.cuda.is_available():
        frames, coords = frames.pin_memory(), coords.pin_memory()
    return frames, coords


def device_labels(coords_dev: torch.Tensor, spec: HierGraphSpec, validate: bool = False):
    """coords int32[B,4,2] on device -> (y, valid) float32[B*N0, 4].  validate=False: no host sync per step (the
    synthetic coordinates are in range by construction; out-of-range ones would NaN-poison the labels)."""
    meta = spec.info()
    y = ops.node_labels(coords_dev, spec.frame_size, meta.level_size, validate=validate)
    y = y.view(-1, y.shape[-1])
    return y, torch.ones_like(y)
