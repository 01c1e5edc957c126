"""pmp_vvc_tip2023_b200 -- B200-native (sm_100a) partition-map prediction for VVC.

Drop-in for the hot path of AolinFeng/PMP-VVC-TIP2023 (Model_QBD / Metrics /
Map2Partition / Inference_QBD): same module names, same ``.pkl`` weights, same
PartitionMat text format; the compute is hand-written CUDA behind the C ABI in
``include/pmp_b200.h`` (``csrc/``), loaded through ctypes by ``_lib``.
"""
__version__ = "0.1.0"
