"""B200-native (sm_100a) degrade -> restore -> classify path behind the reference's PyTorch module contract.

Importable as `b200restore` (see b200restore.py at the repo root; this directory's name is not a Python identifier).

    from b200restore import SimpleUNet, ResUNet, VGG16Judge      # drop-in nn.Modules (same state_dict schema)
    from b200restore import degrade, pipeline                      # fused degradation, end-to-end pipeline
"""
from . import _lib, build, ops, packing, degrade, models, pipeline, synth, generators, imageio, harness, netplan  # noqa: F401
from ._lib import B2RError  # noqa: F401
from .models import ResidualBlock, ResUNet, SimpleUNet, VGG16Judge, vgg16  # noqa: F401
from .pipeline import RestoreClassifyPipeline, all_reduce_counts, shard_range  # noqa: F401
from .generators import CascadeRestorer  # noqa: F401
from .netplan import NetPlan  # noqa: F401

__all__ = ["B2RError", "SimpleUNet", "ResUNet", "ResidualBlock", "VGG16Judge", "vgg16", "RestoreClassifyPipeline",
           "all_reduce_counts", "shard_range", "degrade", "pipeline", "models", "ops", "packing", "synth", "build",
           "generators", "CascadeRestorer", "imageio", "harness", "netplan", "NetPlan"]
