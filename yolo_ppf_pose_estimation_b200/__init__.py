"""B200-native Point-Pair-Feature pose engine (drop-in for the PCL PPF operators the reference's
post-YOLO path is specified against).  Kernels: csrc/ (sm_100a); boundary: include/b200ppf.h.

Python here is host plumbing only (ctypes binding, PCL-shaped mirror classes, synthetic workload
generators, multi-GPU sharding over torch.distributed); all arithmetic runs in libb200ppf.so.
"""
__all__ = ["capi", "build"]
