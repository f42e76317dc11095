"""deepfluids_b200 -- B200 (sm_100a) native implementation of the deep-fluids generator/AE train-step hot path.

Host-side mirror of the reference's Python surface (config / main / Trainer / model / ops) over a C-ABI CUDA
library (`lib/libdeepfluids_b200.so`, declared in include/deepfluids_b200.h).  PyTorch is used only for device
memory, streams and torch.distributed; there is no CPU fallback -- importing `deepfluids_b200.cabi.lib()` without
the built library or without a B200 raises.
"""
__version__ = "0.1.0"
