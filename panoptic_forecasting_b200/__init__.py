"""B200-native bg semantic-forecasting hot path (drop-in for panoptic_forecasting.models on that path).

Host side: Python/PyTorch for device memory and streams; compute: hand-written sm_100a CUDA in
libpf_b200.so reached through the C ABI of include/pf_b200.h.  No CPU fallback."""
__version__ = "0.1.0"
