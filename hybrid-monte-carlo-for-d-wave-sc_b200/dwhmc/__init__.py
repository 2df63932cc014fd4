"""dwhmc -- host side of libdwhmc.so, the B200 implementation of the DwaveHMC.jl molecular-dynamics
force path.  `reference_api` mirrors the reference's single-chain operators; `ChainBatch` is the
batched form used for scans; `parallel` shards chains over the GPUs of a box."""
from ._lib import DwhmcError, EigenConvergenceError, LIB_PATH, version  # noqa: F401
from .batch import ChainBatch, OBS_NAMES, calc_optimal_dt, julia_range, neighbour_tables  # noqa: F401
from .reference_api import (ComputeCache, ModelParameters, ObservablesResult, SimulationState, SpectrumResult,  # noqa: F401
                            build_current_operator, measure_transport_and_spectra, compute_forces, compute_total_energy, diagonalize_H_BdG, hmc_sweep, init_static_H,
                            initialize_cache, initialize_state, measure_observables, refresh_momentum, update_H_BdG)
from .simulation import batch_scan_T, batch_scan_beta, run_simulation, run_simulation_batch, scan_Nt_efficiency  # noqa: E402,F401
