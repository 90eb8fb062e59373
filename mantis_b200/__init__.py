"""mantis_b200 -- B200-native (sm_100a) implementation of the Mantis interleaved multi-image hot path.

Sub-packages:
  csrc/    hand-written CUDA kernels + the C ABI (libmantis_b200.so, declared in include/mantis_b200.h)
  ops      torch-facing wrappers / autograd Functions over the C ABI
  models/  drop-in mirrors of mantis.models.mllava / mantis.models.idefics2 (HF PreTrainedModel API)
"""
__version__ = "0.1.0"
