"""deer_vla_b200 -- B200-native (sm_100a) engine for the DeeR-VLA per-timestep multi-exit forward.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of
``include/deer_engine.h``), the ctypes binding, and a host-side mirror of the reference's call
surface (``MPTFlamingo`` / ``ExitController`` / ``ModelWrapper`` / ``create_model_and_transforms``).
Importing the package does not load the CUDA library; creating an engine does, and fails loudly
if it is missing (there is no CPU fallback).
"""
from .config import DeerConfig, config_3b, config_9b, config_tiny  # noqa: F401

__all__ = ["DeerConfig", "config_3b", "config_9b", "config_tiny"]
