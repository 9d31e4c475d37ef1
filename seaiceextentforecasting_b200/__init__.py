"""seaiceextentforecasting_b200 -- B200 (sm_100a) implementation of the complex-network + GP-regression
sea-ice-extent forecaster's data-parallel core.  Importing the package does not touch the
GPU; the first kernel call loads libsie_b200.so and fails loudly if it (or a CUDA device) is missing."""
from . import config, synthetic  # noqa: F401

__all__ = ["config", "synthetic"]
