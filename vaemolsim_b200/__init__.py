"""vaemolsim_b200 -- B200-native (sm_100a) implementation of the vaemolsim hot path.

Drop-in for the data-parallel path of Monroe-Molecular-Simulation-Group/vae-mol-sim: batched VAE ELBO and
log-probability evaluation (flows, dists, losses, mappings, models, mcmc keep the reference's public API) on top of
hand-written CUDA kernels behind the C ABI of include/vms_b200.h.  No CPU fallback: the first compute call raises if
libvms_b200.so or a CUDA device is missing.
"""
from . import _abi  # noqa: F401
from ._abi import Tensor, as_tensor, synchronize  # noqa: F401
from ._protocols import set_seed  # noqa: F401
from . import dists, flows, losses, mappings, mcmc, models  # noqa: F401

__version__ = '0.1.0'
