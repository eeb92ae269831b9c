"""qclojure_b200 — B200-native state-vector backend behind QClojure's backend protocol.

The product path is `libqcb200.so` (hand-written sm_100a CUDA behind the C ABI in
`include/qcb200.h`); this package is the thin host-side mirror of the reference's
`QuantumBackend` interface plus the ctypes binding.  There is no CPU fallback: creating a
simulator without a CUDA device raises.
"""
__version__ = "0.1.0"
