"""Drop-in for the reference's kurtosis.py — `from kurtosis import KurtosisWeight, RidgeRegularization,
WeightRegularization` (train.py:29): same three class names; KurtosisWeight runs on the fused kernel."""
from bdbnn_b200.losses import KurtosisWeight, RidgeRegularization, WeightRegularization  # noqa: F401
