"""liteasr_b200 -- B200-native (sm_100a) implementation of LiteASR's U2 + hybrid-CTC training hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic runs in
hand-written CUDA kernels behind the C ABI declared in ``include/lasr.h`` (``liblasr.so``).
There is no CPU fallback: ops raise if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
