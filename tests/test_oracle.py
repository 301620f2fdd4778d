"""CPU: the oracle restatement against the golden fixtures made from the live reference module,
and against the live module itself when /root/reference is present."""
import numpy as np
import pytest
import torch

from helpers import EXTRA_IN_NAMES, GOLDEN_NAMES, extra_in_inputs, golden_inputs, load_golden, sha_state
from oracle import cnf_oracle as O


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)  # fixtures were generated single-threaded (fixed GEMM reduction order)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_init_matches_reference_constructor(name):
    g = load_golden(name)
    sd, _, _ = golden_inputs(g)
    assert sha_state(sd) == str(g["weights_sha256"])
    if name == "tiny_shared":
        for k, v in sd.items():
            assert np.array_equal(v.numpy(), g["w:" + k]), k


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_forward_matches_golden(name):
    g = load_golden(name)
    sd, c, l = golden_inputs(g)
    y = O.forward(sd, c, l)
    assert tuple(y.shape) == g["y"].shape
    # bit-identical op order; allow only BLAS-kernel-selection noise between machines
    assert O.rel_l2(y, torch.from_numpy(g["y"])) < 2e-6
    if np.array_equal(y.numpy(), g["y"]):
        return
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=0, atol=5e-6)


@pytest.mark.parametrize("name", [n for n in GOLDEN_NAMES if n != "case1_grid"])
def test_latent_gradient_matches_golden(name):
    g = load_golden(name)
    sd, c, l = golden_inputs(g)
    mask, y_meas = torch.from_numpy(g["mask"]), torch.from_numpy(g["y_meas"])
    loss, _, grad = O.grad_latents(sd, c, l, lambda y: O.sensor_loss(y, y_meas, mask))
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert O.rel_l2(grad.reshape(g["dlatents"].shape), torch.from_numpy(g["dlatents"])) < 1e-4


@pytest.mark.parametrize("name", EXTRA_IN_NAMES)
def test_extra_in_matches_golden(name):
    """a8: SIRENAutodecoder_film_extra_in = the same chain on [extra, coords] (nf_networks.py:503-508)."""
    g = load_golden(name)
    sd, _, lat, cat = extra_in_inputs(g)
    assert sha_state(sd) == str(g["weights_sha256"])
    y = O.forward(sd, cat, lat)
    assert O.rel_l2(y, torch.from_numpy(g["y"])) < 2e-6
    grad = O.grad_latents_from_gout(sd, cat, lat, torch.from_numpy(g["gout"]))
    assert O.rel_l2(grad.reshape(g["dlatents"].shape), torch.from_numpy(g["dlatents"])) < 1e-4


def test_gout_vjp_equals_loss_gradient():
    g = load_golden("case1_shared")
    sd, c, l = golden_inputs(g)
    mask, y_meas = torch.from_numpy(g["mask"]), torch.from_numpy(g["y_meas"])
    y = O.forward(sd, c, l).detach().requires_grad_(True)
    loss = O.sensor_loss(y, y_meas, mask)
    (gout,) = torch.autograd.grad(loss, y)
    vjp = O.grad_latents_from_gout(sd, c, l, gout)
    assert O.rel_l2(vjp.reshape(g["dlatents"].shape), torch.from_numpy(g["dlatents"])) < 1e-4


def test_fp64_noise_floor():
    g = load_golden("case1_shared")
    sd, c, l = golden_inputs(g)
    y64 = O.forward(O.to_dtype(sd, torch.float64), c.double(), l.double())
    assert O.rel_l2(torch.from_numpy(g["y"]), y64) < 5e-6


def test_live_reference_bit_identical():
    Ref = O.load_reference_module()
    if Ref is None:
        pytest.skip("/root/reference not present (GPU box)")
    for name, (cin, L, cout, nl, H) in O.CASE_SHAPES.items():
        torch.manual_seed(0)
        m = Ref(cin, L, cout, nl, H)
        sd = O.init_params(cin, L, cout, nl, H, seed=0)
        assert list(sd.keys()) == list(m.state_dict().keys())
        for k, v in m.state_dict().items():
            assert torch.equal(sd[k], v), (name, k)
        c, l = O.synthetic_inputs(cin, L, 2, 97)
        assert torch.equal(O.forward(sd, c[None], l[:, None]), m(c[None], l[:, None])), name
