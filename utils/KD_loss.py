"""Drop-in for the reference's utils/KD_loss.py (imported at train.py:34): same class names."""
from bdbnn_b200.losses import DistributionLoss, DistributionLoss_layer  # noqa: F401
