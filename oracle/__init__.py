"""TEST INFRASTRUCTURE: ctypes binding of the CPU oracle (oracle/cubep3m_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
from .binding import Oracle, build_oracle, oracle_fft3d  # noqa: F401
