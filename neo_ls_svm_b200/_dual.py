"""Dual solve orchestration: `NeoLSSVM._optimize_α̂_γ` and the dual predict branches behind the C ABI.

Mirrors /root/reference/src/neo_ls_svm/_neo_ls_svm.py:191-325 (ρ = 1), :473-475 and :668-671.
Single GPU ("replicas only"): the n×n kernel matrix, its eigendecomposition and the LOO
sweep over 128 γ all run on the device; the n×G×n tensor of the reference is never formed.
"""

from __future__ import annotations

import numpy as np

from ._primal import gamma_grid, select_gamma

N_GAMMAS_DUAL = 128  # _neo_ls_svm.py:270
_DEVICE_STATE = "_device_state"


def _clip_correct_side(res, y):
    res[(y > 0) & (res > 0)] = 0
    res[(y < 0) & (res < 0)] = 0


def fit_into(model, Xt: np.ndarray, y: np.ndarray, s: np.ndarray):
    """Solve the dual system for the transformed rows Xt and store the fitted attributes on `model`."""
    from sklearn.metrics import accuracy_score, r2_score

    ctx, torch, dev = model._gpu()
    dt = Xt.dtype
    classifier = model._estimator_type == "classifier"
    y64 = np.asarray(y, dtype=np.float64)
    s_norm = np.asarray(s, dtype=np.float64)
    s_norm = s_norm / np.sum(s_norm)  # :252
    sn = s_norm / np.median(np.abs(s_norm))  # :253
    gammas_np = gamma_grid(N_GAMMAS_DUAL)

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)

    Xd, yd, sd, snd = up(Xt), up(y64), up(s_norm), up(sn)
    sums, yhat_loo, lam = ctx.dual_sweep(Xd, yd, sd, snd, up(gammas_np), classifier)
    sums_np = sums.cpu().numpy()
    opt, _ = select_gamma(sums_np, classifier)
    gamma = float(gammas_np[opt])
    fin = ctx.dual_finalize(len(y64), yd, snd, gamma, cholesky=True)
    yhat_opt = yhat_loo[:, opt].cpu().numpy()
    loo = yhat_opt - y64  # :287
    if classifier:
        _clip_correct_side(loo, y64)
    model.γs_ = gammas_np.astype(dt)
    model.loo_errors_γs_ = sums_np[0].astype(dt)
    model.loo_residuals_ = loo.astype(dt)
    model.loo_ŷ_ = (y64 + loo).astype(dt)
    model.loo_error_ = model.loo_errors_γs_[opt]
    if classifier:
        model.loo_score_ = accuracy_score(y64, np.sign(yhat_opt), sample_weight=s_norm)
    else:
        model.loo_score_ = r2_score(y64, yhat_opt, sample_weight=s_norm)
    model.L_ = (fin["U"].cpu().numpy().astype(dt), False)  # cho_factor layout (:313)
    res = fin["Falpha"].cpu().numpy() - y64  # :315
    if classifier:
        _clip_correct_side(res, y64)
    model.residuals_ = res.astype(dt)
    model.loo_std_ = np.sqrt(fin["sigma2"].cpu().numpy()).astype(dt)  # :321-323
    alpha = fin["alpha"]
    model.__dict__[_DEVICE_STATE] = {
        "kind": "dual", "Xt": Xd, "alpha": alpha, "alpha_sum": float(alpha.sum()), "Bt": fin["Bt"], "w": fin["w"],
    }
    return alpha.cpu().numpy().astype(dt), model.γs_[opt]


def _device_state(model):
    """Device copies of what the dual predict needs (rebuilt from host attributes after unpickling)."""
    st = model.__dict__.get(_DEVICE_STATE)
    if st is not None and st.get("kind") == "dual":
        return st
    ctx, torch, dev = model._gpu()
    n = model.X_.shape[0]
    U = torch.from_numpy(np.ascontiguousarray(model.L_[0], dtype=np.float64)).to(dev)
    # (γS⁻² + F)⁻¹ = U⁻¹ U⁻ᵀ  ⇒  σ² = 1 − ‖(K U⁻¹)ᵢ‖²: Bt = (U⁻¹)ᵀ with unit weights (nls_triangular_inverse reads
    # the upper triangle only, like cho_solve does).
    Uinv = ctx.triangular_inverse(U)
    alpha = torch.from_numpy(np.asarray(model.α̂_, dtype=np.float64)).to(dev)
    st = {
        "kind": "dual",
        "Xt": torch.from_numpy(np.ascontiguousarray(model.X_, dtype=np.float64)).to(dev),
        "alpha": alpha, "alpha_sum": float(np.sum(np.asarray(model.α̂_, dtype=np.float64))),
        "Bt": Uinv.T.contiguous(), "w": torch.ones(n, dtype=torch.float64, device=dev),
    }
    model.__dict__[_DEVICE_STATE] = st
    return st


def predict(model, X: np.ndarray, want_decision: bool, want_std: bool):
    """ŷ = K(x, X_) α̂ + Σα̂ and σ = sqrt(1 − kᵀ(γS⁻² + F)⁻¹k) for raw input rows X (device tensors)."""
    ctx, torch, dev = model._gpu()
    st = _device_state(model)
    if "shift" not in st:
        shift, W = model.dual_feature_map_.device_weights(model.n_features_in_)
        st["shift"], st["W"] = torch.from_numpy(shift).to(dev), torch.from_numpy(W).to(dev)
    Xraw = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(dev)
    Xq = ctx.affine_map(Xraw, st["shift"], st["W"])  # :473, :668
    aff = model.dual_feature_map_
    if aff.append_features and getattr(aff, "A_", aff.A) is not None:
        Xq = torch.cat([Xraw, Xq], dim=1).contiguous()  # transform() hstacks [X | Z] (_affine_feature_map.py:90-91)
    return ctx.dual_predict(
        Xq, st["Xt"], alpha=st["alpha"] if want_decision else None, alpha_sum=st["alpha_sum"],
        Bt=st["Bt"] if want_std else None, w=st["w"] if want_std else None, want_std=want_std,
    )
