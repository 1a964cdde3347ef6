"""prmers_b200 -- B200-native (sm_100a) implementation of PrMers' Marin modular-squaring engine.

Only the hot path: the register engine behind `engine::Reg` (include/marin/engine.h) as hand-written
CUDA kernels behind a C ABI (include/marin_b200.h), plus this thin host-side mirror of the API.
"""
from .engine import Engine, Digit, MarinError, create_gpu, transform_size, load_library, exported_symbols  # noqa: F401

__version__ = "0.1.0"
