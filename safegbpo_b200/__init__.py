"""safegbpo_b200 -- B200-native safeguarded rollout (drop-in for TimWalter/SafeGBPO's hot path).

The compute path is hand-written sm_100a CUDA behind the C ABI of include/sgbpo.h; importing the
package does not load it, the first op does, and raises if it is missing (there is no fallback).
"""
__version__ = "0.1.0"
