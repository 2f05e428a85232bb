"""cubep3m_b200 — B200-native `particle_mesh` for CUBEP3M behind a C ABI (include/cubep3m_b200.h).

This package is the thin Python harness: ctypes binding of the CUDA library, the kernel tables, the
synthetic IC generator and the P(k) twin.  The product is cubep3m_b200/csrc (CUDA + C ABI)."""
from .abi import Config, StepOut, Clock, CheckpointHeader, checkpoint_name, default_config, copy_config, max_np, ST_NAMES  # noqa: F401
