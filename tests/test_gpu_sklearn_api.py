"""sklearn conformance of the estimator on the GPU path (the reference's own tests run check_estimator,
tests/test_neo_ls_svm.py:111-116).  Under the installed sklearn the reference itself fails four checks
(SURVEY.md §4); the new class must pass at least the same set."""

import numpy as np
import pytest

from neo_ls_svm_b200 import NeoLSSVM

pytestmark = pytest.mark.gpu

# Checks the unmodified reference fails under sklearn 1.9 — tolerated here as well.
KNOWN_REFERENCE_FAILURES = {
    "check_estimator_tags_renamed",
    "check_n_features_in_after_fitting",
    "check_all_zero_sample_weights_error",
    "check_sample_weight_equivalence_on_dense_data",
}


@pytest.mark.parametrize("kind", ["regressor", "classifier"])
def test_check_estimator(kind):
    from sklearn.utils.estimator_checks import check_estimator

    results = check_estimator(NeoLSSVM(estimator_type=kind), on_fail=None)
    failed = {r["check_name"] for r in results if r["status"] == "failed"}
    passed = [r for r in results if r["status"] == "passed"]
    unexpected = failed - KNOWN_REFERENCE_FAILURES
    assert not unexpected, {r["check_name"]: str(r["exception"])[:300] for r in results if r["check_name"] in unexpected}
    assert len(passed) >= 40


def test_clone_and_get_params():
    from sklearn.base import clone

    est = NeoLSSVM(dual=False, random_state=7)
    twin = clone(est)
    assert twin.get_params()["random_state"] == 7 and twin.get_params()["dual"] is False
    rng = np.random.default_rng(0)
    X = rng.standard_normal((1500, 4))
    y = X[:, 0] + 0.1 * rng.standard_normal(1500)
    twin.fit(X, y)
    assert not hasattr(est, "β̂_") and hasattr(twin, "β̂_")
