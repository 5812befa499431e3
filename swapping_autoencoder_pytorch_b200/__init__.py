"""B200-native (sm_100a) implementation of the Swapping-Autoencoder training hot path.

Layout:
  csrc/ + libsae_b200.so   hand-written CUDA kernels behind the C ABI of include/sae_b200.h
  _lib.py / backend.py     ctypes binding; tensor-level entry points (no torch types cross the ABI)
  stylegan2_op/            drop-in for reference models/networks/stylegan2_op (upfirdn2d, fused_leaky_relu, ...), the conv
                           Functions (conv.py) and the block-level autograd nodes (blocks.py)
  stylegan2_layers.py      drop-in for reference models/networks/stylegan2_layers.py (the operator surface)
  networks/                E, G, D, Dpatch with the reference's state_dict keys
  model.py / optimizer.py  loss graph and D/G/R1 training driver (restated callers)
  graphs.py                CUDA-graph capture / replay of the D, G and R1 half-steps (opt.cuda_graphs)
  parallel.py              one-process-per-GPU NCCL data parallelism behind the MultiGPUModelWrapper surface
  options.py               the reference's default option set as a Namespace
"""
from . import _lib, backend  # noqa: F401
from .options import default_options  # noqa: F401

__all__ = ["default_options", "backend", "build_library", "create_model", "create_optimizer"]


def build_library(verbose=False):
    return _lib.build_library(verbose=verbose)


def create_model(opt):
    """reference models/__init__.py:57-72: instantiate, initialize, wrap."""
    from .model import SwappingAutoencoderModel
    from .parallel import MultiGPUModelWrapper
    instance = SwappingAutoencoderModel(opt)
    instance.initialize()
    return MultiGPUModelWrapper(opt, instance)


def create_optimizer(opt, model):
    from .optimizer import SwappingAutoencoderOptimizer
    return SwappingAutoencoderOptimizer(model)
