"""scone_gcn_b200 — B200-native (sm_100a) implementation of the SCoNe training / inference hot path.

Host side mirrors the reference's Python surface (nglaze00/SCoNe_GCN):
  scone_gcn_b200.trajectory_experiments   hyperparams, scone_func / ebli_func / bunch_func, data_setup, train_model
  scone_gcn_b200.scone_trajectory_model   Scone_GCN
  scone_gcn_b200.synthetic_data_gen       dataset folder format, generators
All arithmetic of the hot path runs in hand-written CUDA kernels behind the C ABI in include/scone_b200.h
(libscone_b200.so).  There is no CPU fallback: importing the binding without the built library, or
calling it without a CUDA device, raises.
"""
from ._lib import lib, library_path, LibraryMissing  # noqa: F401
from .complex import SimplicialComplex, flows_to_csr  # noqa: F401
from .model import SconeModel  # noqa: F401

__version__ = '0.1.0'
