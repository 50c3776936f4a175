"""
odil_b200 -- B200-native (sm_100a) residual-and-gradient engine behind the ODIL Python API.
Host-side mirror of cselab/odil's Domain / operator(ctx) / Problem / optimizer surface on top of
hand-written CUDA kernels reached through a ctypes C ABI (include/odil_b200.h).
"""
__version__ = "0.1.0"
