"""openfly_b200 -- B200-native (sm_100a) force-evaluation path for openFLY behind a C ABI.

`capi`    ctypes binding of include/openfly_b200.h (the drop-in boundary)
`fly`     host-side mirror of the reference's classes (Generic / List / LBFGS / Dimer) over that ABI
`inputs`  synthetic systems and start states of the BASELINE configurations (numpy, set-up only)
`shard`   replica sharding across the GPUs of a box and the host-side merge of the finished records
"""
from . import capi, fly  # noqa: F401

__all__ = ["capi", "fly"]
