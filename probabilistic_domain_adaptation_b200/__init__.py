"""B200-native (sm_100a) Probabilistic U-Net training / Monte-Carlo inference path.

Drop-in for `prob_utils.my_models` of computational-cell-analytics/Probabilistic-Domain-Adaptation:
same module API and state_dict, arithmetic in hand-written CUDA behind a C ABI (include/pda_b200.h).
"""
from .my_models import ProbabilisticUnet, l2_regularisation  # noqa: F401

__all__ = ["ProbabilisticUnet", "l2_regularisation"]
