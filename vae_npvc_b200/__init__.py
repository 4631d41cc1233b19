"""B200-native ConvVAE hot path of JeremyCCHsu/vae-npvc: hand-written sm_100a CUDA behind a C-ABI
(``include/npvc_b200.h``), bound with ctypes; PyTorch only owns memory, streams and NCCL."""
from .arch import vcc2016_vae_arch  # noqa: F401
