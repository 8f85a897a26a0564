"""CPU oracle of the vksift detect + match path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (vulkansift_b200) never does.
"""
from .oracle import (Oracle, OracleConfig, FEATURE_DTYPE, MATCH_DTYPE, build, lib_path, match_descriptors,
                     match_features, arith, seed_image, downsample_nearest)  # noqa: F401
