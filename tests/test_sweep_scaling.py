"""CPU model of the INT8 γ sweep's operand scaling (scripts/sweep_int8_study.py: operands quantised to 56 fractional bits
relative to their row maxima as csrc/ozaki.cuh does; the kernel's integer plane products are exact).  Pins the two facts the
kernel design rests on: a plain fixed-point split of P, U and rγ is NOT accurate enough on ill-conditioned fits, and the
two-sided column scaling d_k = |λ_k| + γ_ref brings everything to FP64 rounding level with a single γ group already."""

import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _study():
    spec = importlib.util.spec_from_file_location("sweep_int8_study", os.path.join(ROOT, "scripts", "sweep_int8_study.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.study


def test_two_sided_scaling_is_fp64_accurate_and_plain_split_is_not():
    res = _study()(["reg_small", "ragged:130:3:40"], group_counts=(0, 1, 2), verbose=False)
    for name, by_groups in res.items():
        for groups in (1, 2):
            r = by_groups[groups]
            assert r["opt_equal"], (name, groups)
            assert r["curve"] < 1e-14 and r["num"] < 1e-14 and r["den"] < 1e-14 and r["loo"] < 1e-13, (name, groups, r)
            assert r["elementwise"] < 0.01, (name, groups, r)
    # the ill-conditioned shape (λ_min = 6e-9 below γ_min = 1e-6) is where the plain split breaks the elementwise bar
    plain = res["ragged:130:3:40"][0]
    assert plain["num"] > 1e-11 and plain["elementwise"] > 1.0, plain
