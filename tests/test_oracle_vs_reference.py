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
    from oracle import make_golden as mg

    mg.OUT = str(tmp_path)
    RN = mg.ref_networks.RunningNorm
    mg.disc_case("disc_gail_hc", "gail", 17, 6, False, dict(normalize_input_layer=RN), 64, 64, 4, 0)
    mg.disc_case("disc_airl_nonorm", "airl", 6, 2, False,
                 dict(reward_hid_sizes=(32, 32), potential_hid_sizes=(32,)), 16, 16, 3, 6, shaped=True)
    mg.running_norm_case()
    mg.buffer_case()
    mg.rollout_case()
    mg.expert_loader_case()
    for name in ("disc_gail_hc", "disc_airl_nonorm", "running_norm", "replay_buffer", "rollout_order",
                 "expert_loader"):
        new = np.load(os.path.join(str(tmp_path), name + ".npz"), allow_pickle=True)
        old = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
        assert set(new.files) == set(old.files)
        for k in old.files:
            if old[k].dtype.kind in "fc":
                np.testing.assert_allclose(new[k], old[k], rtol=1e-6, atol=1e-7, err_msg=f"{name}:{k}")
            else:
                np.testing.assert_array_equal(new[k], old[k], err_msg=f"{name}:{k}")


def test_reference_modules_used_are_the_real_ones():
    im = refimport.load()
    assert im.__file__.startswith("/root/reference/src/imitation")
    from imitation.algorithms.adversarial import common

    assert common.__file__ == "/root/reference/src/imitation/algorithms/adversarial/common.py"
