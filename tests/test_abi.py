"""CPU: the C-ABI library loads, exports every symbol include/visgeom_b200.h declares, and has no
CPU fallback (compute entry points fail loudly without a device)."""
import os
import re

import numpy as np
import pytest

import synthdata as sd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "visgeom_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(vg_[a-z_0-9]+)\s*\(", text)
    return sorted(set(n for n in names if n != "vg_allreduce_fn"))


def test_every_declared_symbol_is_exported(vg):
    lib = vg.lib()
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/visgeom_b200.h but not exported: {missing}"


def test_model_tables(vg):
    assert [vg.lib().vg_model_num_params(m) for m in (0, 1, 2)] == [6, 5, 10]
    assert vg.lib().vg_model_num_params(9) < 0 and b"invalid camera model name" in vg.lib().vg_last_error()
    assert vg.model_bounds(sd.EUCM)[:2] == [(0.0, 1.0), (0.1, 10.0)]       # eucm.h:228-246
    assert vg.model_bounds(sd.UCM)[0] == (0.0, 3.0)                          # ucm.h:199-215
    assert vg.model_bounds(sd.MEI)[1] == (-10.0, 10.0) and vg.model_bounds(sd.MEI)[9] == (1.0, 1e5)   # mei.h:287-313
    assert vg.hessian_entries(sd.EUCM, 1) == 91 and vg.hessian_entries(sd.MEI, 1) == 17 * 18 // 2
    assert vg.hessian_entries(sd.EUCM, 2) == 19 * 20 // 2


def test_no_cpu_fallback(vg):
    if vg.device_count() > 0:
        pytest.skip("a GPU is present; the loud-failure path is exercised on CPU-only hosts")
    d = sd.make_mono(sd.EUCM, 2, seed=3)
    with pytest.raises(vg.VisgeomError, match="no CUDA device"):
        vg.eval_chain(sd.EUCM, d["intr_init"], d["board"], d["obs"], [d["xi_init"]], [0], [0])
    with pytest.raises(vg.VisgeomError, match="no CUDA device"):
        vg.Problem()
    # the prior functors and visualCov have no CPU path either
    x = np.zeros((2, 6)); x[:, 2] = 1.0
    with pytest.raises(vg.VisgeomError, match="no CUDA device"):
        vg.eval_transformation_prior(np.ones((2, 6)), x, x)
    with pytest.raises(vg.VisgeomError, match="no CUDA device"):
        vg.eval_odometry_prior(0.1, 0.1, 0.01, x, x, x, x)
    with pytest.raises(vg.VisgeomError, match="no CUDA device"):
        vg.visual_cov(sd.EUCM, d["intr_init"], x[0], d["board"], 0.25, x)


def test_product_does_not_touch_the_oracle():
    """Nothing under visgeom_b200/ or include/ may reference oracle/ (the checker is not the product)."""
    bad = []
    for base in ("visgeom_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"oracle|vgo_|vgref_", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
