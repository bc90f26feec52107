"""mcrg_b200 — B200-native (sm_100a) implementation of the MCRG hot path of kim-jane/MCRG:
checkerboard Metropolis sweeps over bit-packed lattices, b=2 majority-rule block-spin pyramid, and the
correlator / cross-correlator accumulation that feeds the linearised RG matrix.

Layout:  csrc/  CUDA kernels + the C ABI (include/mcrg_b200.h)   ->  libmcrg_b200.so
         host/  C++ drop-in classes with the reference's Lattice / IsingModel / MonteCarloRenormalizationGroup API
         capi.py / analysis.py / dist.py   Python plumbing used by the tests and bench.py
"""
from . import analysis, capi, dist  # noqa: F401
from .capi import Context, McrgError  # noqa: F401
