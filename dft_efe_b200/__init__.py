"""B200-native H·X / Chebyshev-filter / Rayleigh-Ritz hot path for dft-efe.

The product is the C-ABI shared library built from ``csrc/`` (hand-written CUDA
for sm_100a) plus the C++ header mirror of the reference operator API in
``include/``.  The Python in this package is harness: a synthetic mesh
generator (``synth``) and a ctypes binding (``capi``) used by tests and bench.
"""
