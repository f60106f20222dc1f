"""jams_b200 — B200-native llg-heun + exchange hot path for stonerlab/jams.

The numerical work lives in ``libjams_b200.so`` (hand-written sm_100a kernels behind the C ABI of
``include/jams_b200.h``); this package is the host-side mirror of the JAMS plugin surface.
"""
from . import capi, consts, lattice  # noqa: F401
from .capi import Context, JamsB200Error  # noqa: F401
from .lattice import Lattice, Material  # noqa: F401
from .solver import (AppliedFieldHamiltonian, B200HeunLLGSolver, EnergyMonitor, ExchangeHamiltonian,  # noqa: F401
                     MagnetisationMonitor, UniaxialAnisotropyHamiltonian, ZeemanHamiltonian, create_hamiltonian,
                     create_solver)

__version__ = "0.1.0"
