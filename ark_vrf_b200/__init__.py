"""ark-vrf_b200: B200-native (sm_100a) batch verification of ark-vrf's Thin VRF.

The product is `libavrf_gpu.so` (C ABI in include/avrf.h); this package is the Python
mirror of the reference's `thin::BatchVerifier` / `thin::Verifier` interface above it,
plus the feeder operations and the synthetic-workload generator used by tests and bench.
"""
from ._lib import AvrfError, LIB_PATH, load  # noqa: F401
from .thin import (BatchItem, BatchServer, BatchVerifier, Error, Format, HashPool, InvalidData, Proof, Public, ShardedBatchVerifier,  # noqa: F401
                   Suite, Tap, VerificationFailure, combine_partials, init_multi, seed_of_stream)

__all__ = ["AvrfError", "BatchItem", "BatchServer", "BatchVerifier", "Error", "Format", "HashPool", "InvalidData", "Proof", "Public", "ShardedBatchVerifier", "Suite", "init_multi",
           "Tap", "VerificationFailure", "combine_partials", "seed_of_stream", "load", "LIB_PATH"]
