"""mmc-b200: B200-native (sm_100a CUDA, FP64) energy engine for the Metropolis Monte Carlo
hot path of BradenDKelly/MetropolisMonteCarlo — see DESIGN.md and include/mmc_b200.h."""
__version__ = "0.1.0"
