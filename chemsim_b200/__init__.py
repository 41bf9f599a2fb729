"""chemsim_b200 — B200-native D2Q9 collide+stream path of taktoa/chemsim's lbm.rs.

`chemsim_b200.lbm` mirrors the reference's `lbm` module over the C ABI of
include/chemsim_lbm.h (libchemsim_lbm.so, built by `python -m chemsim_b200.build`);
`chemsim_b200.scenarios` holds the initial conditions main.rs and BASELINE.json use.
Importing the package does not load the CUDA library; the first use of
`chemsim_b200.lbm.State` does, and fails loudly if it has not been built.
"""
__version__ = "0.1.0"
__all__ = ["lbm", "scenarios", "build"]
