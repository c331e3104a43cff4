"""Build-container only: regenerate the golden vectors from the REAL reference
(/root/reference/src, imported through oracle/_shim) and check they equal the committed
fixtures -- i.e. tests/golden/*.npz really are outputs of the reference's own modules.
Skipped where /root/reference does not exist (the GPU box)."""
import os

import numpy as np
import pytest

from oracle import refimport

pytestmark = [pytest.mark.refsrc,
              pytest.mark.skipif(not refimport.available(), reason="/root/reference not present")]

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_goldens_reproduce_from_reference(tmp_path):
    """EVERY file under tests/golden/ is regenerated from the reference's own modules and compared with the committed
    fixture (no golden is taken on trust)."""
    from oracle import make_golden as mg

    mg.OUT = str(tmp_path)
    cases = mg.all_cases()
    committed = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith("demo_"))
    assert sorted(n for n, _ in cases) == committed, "a golden file without a generator (or the reverse)"
    for name, thunk in cases:
        thunk()
        new = np.load(os.path.join(str(tmp_path), name + ".npz"), allow_pickle=True)
        old = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
        assert set(new.files) == set(old.files), name
        for k in old.files:
            if old[k].dtype.kind in "fc":
                np.testing.assert_allclose(new[k], old[k], rtol=1e-6, atol=1e-7, err_msg=f"{name}:{k}")
            elif old[k].dtype.kind == "O":
                assert [str(x) for x in np.ravel(new[k])] == [str(x) for x in np.ravel(old[k])], f"{name}:{k}"
            else:
                np.testing.assert_array_equal(new[k], old[k], err_msg=f"{name}:{k}")


def test_demo_fixtures_are_cut_from_the_reference_rollouts(tmp_path):
    """tests/golden/demo_*.npz = leading trajectories of the reference's own expert rollouts (oracle/make_demo_fixture.py)."""
    from oracle import make_demo_fixture as mdf

    for src, name, k in (("cartpole_0/rollouts/final.npz", "demo_cartpole_legacy", 4),
                         ("pendulum_0/rollouts/final.npz", "demo_pendulum_legacy", 3)):
        out = os.path.join(str(tmp_path), name + ".npz")
        mdf.cut(os.path.join(mdf.REF, src), out, k)
        new, old = np.load(out, allow_pickle=True), np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
        assert set(new.files) == set(old.files)
        for key in old.files:
            if old[key].dtype.kind == "O":
                assert [str(x) for x in new[key]] == [str(x) for x in old[key]]
            else:
                np.testing.assert_array_equal(new[key], old[key], err_msg=f"{name}:{key}")


def test_reference_modules_used_are_the_real_ones():
    im = refimport.load()
    assert im.__file__.startswith("/root/reference/src/imitation")
    from imitation.algorithms.adversarial import common

    assert common.__file__ == "/root/reference/src/imitation/algorithms/adversarial/common.py"
