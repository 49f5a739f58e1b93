"""risc0_b200 — B200-native (sm_100a) backend for the RISC Zero STARK prover hot path.

The product is `lib/libr0b200.so` (hand-written CUDA behind the C ABI of include/r0b200.h). This package is the
host-side mirror of the reference's `risc0_zkp::hal::Hal` trait over that ABI (`B200Hal`, `Buffer`), used by the parity
tests and bench.py. There is NO CPU fallback: without the built library or without a CUDA device the calls raise.
"""
from ._lib import LIB_PATH, R0B200Error, lib_available, load_library  # noqa: F401
from .hal import B200Hal, Buffer, DeviceTrace, Rv32imCircuitHal, SegmentProver, WitnessGenerator  # noqa: F401
