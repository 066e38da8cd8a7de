"""B200-native Neo LS-SVM hot path behind the reference's sklearn-compatible API."""

from ._affine import AffineFeatureMap, AffineNormalizer, AffineSeparator
from ._clqr import CoherentLinearQuantileRegressor
from ._feature_maps import (
    KernelApproximatingFeatureMap,
    OrthogonalRandomFourierFeatures,
    RandomFourierFeatures,
)
from ._multi import devices, set_devices
from ._neo_ls_svm import NeoLSSVM

__all__ = [
    "NeoLSSVM",
    "AffineFeatureMap",
    "AffineNormalizer",
    "AffineSeparator",
    "CoherentLinearQuantileRegressor",
    "KernelApproximatingFeatureMap",
    "OrthogonalRandomFourierFeatures",
    "RandomFourierFeatures",
    "devices",
    "set_devices",
]
