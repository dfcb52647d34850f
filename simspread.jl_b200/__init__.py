"""simspread_b200 -- B200-native implementation of SimSpread.jl's resource-spreading hot path.

The directory is called `simspread.jl_b200` (not an importable identifier); import it through the
`simspread_b200` shim at the repository root.  Layout:

  csrc/     hand-written sm_100a CUDA kernels + the C ABI (include/simspread_b200.h)
  lib/      the built libsimspread_b200.so (git-ignored, travels with the snapshot)
  host.py   Python mirror of the reference's exported Julia API over that ABI
  julia/    the same host layer in Julia (`ccall`), for SimSpread.jl users
"""
from ._lib import SimSpreadError, header_symbols, lib  # noqa: F401
from .namedarray import NamedArray  # noqa: F401
from .host import (  # noqa: F401
    AuPRC, AuROC, BEDROC, Context, alpha_sweep, cross_validate, DCsr, DIVec, DMat, Graph, accuracy, balancedaccuracy, clean_, construct, cutoff,
    cutoff_, f1score, featurize, featurize_, jaccard_featurize, tanimoto_featurize_bits, k, mcc, maxperformance, merge_cross_validation, meanperformance, meanstdperformance, precision, precisionatL, predict, read_namedmatrix, recall,
    recallatL, recommend_topl, save, split, spread, validity_ratio, writedlm,
)
from ._build import build, lib_path  # noqa: F401

__version__ = "0.1.0"
