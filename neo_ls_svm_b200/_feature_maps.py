"""Kernel-approximating random feature maps φ: Rᵈ → C^{D+1}.

Mirrors the reference's transformers (/root/reference/src/neo_ls_svm/_feature_maps.py):
`KernelApproximatingFeatureMap` (:58-114), `RandomFourierFeatures` (:117-203) and
`OrthogonalRandomFourierFeatures` (:206-223).

* `fit` stays on the host: the affine pre-pass and the random frequencies Z_ are drawn with the same
  NumPy RandomState call sequence as the reference, so W = A_/scaleᵀ is identical for identical seeds.
* `transform` is stage 1 of the hot path: z = (x - shift) W and φ = [exp(-i z)/√D | 1] are computed by
  the sm_100a feature-map kernel (`nls_feature_map`); the result is returned as a host complex array
  for API parity.  The solver itself never materialises φ (see `_primal.py`).
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from functools import cached_property

import numpy as np
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.utils import check_array, check_random_state

from ._affine import AffineFeatureMap, AffineSeparator


class KernelApproximatingFeatureMap(ABC, BaseEstimator, TransformerMixin):
    """Abstract kernel-approximating feature map (reference :58-114)."""

    def __init__(self, affine_feature_map: AffineFeatureMap | None = None, num_features: int = 512,
                 random_state: int | np.random.RandomState | None = 42):
        self.num_features, self.D = num_features, num_features
        self.affine_feature_map = affine_feature_map or AffineSeparator()
        self.random_state = random_state

    @cached_property
    @abstractmethod
    def complexity_matrix(self):
        """Regulariser C of the primal objective γ β̂ᴴ C β̂ (identity for all shipped maps)."""

    @abstractmethod
    def fit(self, X, y=None, sample_weight=None):
        self.affine_feature_map.fit(X, y, sample_weight)
        self.n_features_in_ = X.shape[1]
        return self

    @abstractmethod
    def transform(self, X):
        ...


class RandomFourierFeatures(KernelApproximatingFeatureMap):
    """Random Fourier Features for the Gaussian kernel exp(-||A(x-y)||²/2)."""

    @classmethod
    def _fourier_features(cls, d: int, D: int, dtype, random_state):
        rng = check_random_state(random_state)
        return rng.randn(d, D).astype(dtype)

    @cached_property
    def complexity_matrix(self):
        # The reference hard-codes the fast diagonal approximation (:134, :43-45): C = I_{D+1}.
        return np.eye(self.D + 1, dtype=self.Z_.dtype)

    def fit(self, X, y=None, sample_weight=None):
        super().fit(X, y, sample_weight)
        aff = self.affine_feature_map
        A = getattr(aff, "A_", aff.A)
        width = A.shape[1] if A is not None else X.shape[1]
        self.Z_ = self._fourier_features(width, self.D, X.dtype, self.random_state)
        # Fold the frequencies into the affine map: z = (x - shift) diag(1/scale) A Z_.
        aff.A_ = A @ self.Z_ if A is not None else self.Z_
        return self

    def device_weights(self, n_features: int):
        """(shift, W) as float64 host arrays with z = (x - shift) W."""
        return self.affine_feature_map.device_weights(n_features)

    def transform(self, X):
        """φ(X) = [exp(-i z)/√D | 1] ∈ C^{n×(D+1)}, computed on the GPU (stage 1)."""
        import torch

        from . import _lib

        X = check_array(X)
        out_dtype = np.complex64 if X.dtype == np.float32 else np.complex128
        shift, W = self.device_weights(X.shape[1])
        ctx = _lib.context()
        dev = torch.device("cuda", ctx.device)
        Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(dev)
        phi = ctx.feature_map(Xd, torch.from_numpy(shift).to(dev), torch.from_numpy(W).to(dev))
        return phi.cpu().numpy().astype(out_dtype, copy=False)


class OrthogonalRandomFourierFeatures(RandomFourierFeatures):
    """Orthogonal Random Features: per-block orthonormalised frequencies with χ-distributed norms."""

    @classmethod
    def _fourier_features(cls, d: int, D: int, dtype, random_state):
        rng = check_random_state(random_state)
        Z = rng.randn(d, D).astype(dtype)
        for j0 in range(0, D, d):  # orthonormalise each block of d columns (:216-218)
            Z[:, j0 : j0 + d] = np.linalg.qr(Z[:, j0 : j0 + d])[0]
        # Restore the norm distribution of Gaussian vectors (:220-221).
        Z *= np.sqrt(rng.chisquare(d, size=(1, Z.shape[1])).astype(dtype))
        return Z
