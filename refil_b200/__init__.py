"""refil_b200 -- B200-native (sm_100a) hot path of REFIL: Group Matching env kernel + fused learner kernels behind a
C ABI (include/refil_b200.h), with host classes mirroring the reference's controllers / learners / envs API."""
__version__ = "0.1.0"
