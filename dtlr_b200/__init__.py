"""dtlr_b200 -- B200-native (sm_100a) implementation of the DTLR DINO-DETR hot path.

Host side: Python over PyTorch tensors (device memory/streams only), mirroring the reference's
models.dino API (build_dino / DINO.forward / SetCriterion.loss_CTC / PostProcess).
Device side: hand-written CUDA kernels behind the C-ABI in include/dtlr_b200.h (libdtlr_b200.so).
There is no CPU fallback: every op raises if the extension is missing or the tensors are not CUDA.
"""
__version__ = "0.1.0"
