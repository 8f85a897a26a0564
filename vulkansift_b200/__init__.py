"""vulkansift_b200 -- B200-native SIFT detect + 2-NN match behind the vksift_* C ABI.

The product is the shared library vulkansift_b200/lib/libvulkansift.so (CUDA,
sm_100a).  This package is its thin ctypes binding plus the synthetic workload
generators used by tests and bench.py.  There is no CPU fallback: importing
`vulkansift_b200.api` fails loudly if the library has not been built.
"""
__version__ = "0.1"
