"""CPU: host-side logic of the drop-in module (no kernels are launched here)."""
import torch
import pytest

import confild_b200 as cb
from confild_b200.nf_networks import canonicalize
from oracle import cnf_oracle as O


def test_state_dict_layout_and_init_match_reference_order():
    torch.manual_seed(0)
    m = cb.SIRENAutodecoder_film(2, 128, 3, 10, 128)
    sd = O.init_params(2, 128, 3, 10, 128, seed=0)
    assert list(m.state_dict().keys()) == list(sd.keys())
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    assert m.nl.w0 == 30.0
    assert len(m.net1) == 12 and len(m.net2) == 11


def test_positional_and_keyword_construction():
    a = cb.SIRENAutodecoder_film(3, 384, 3, 15, 384)  # measurements.py:207 style
    b = cb.SIRENAutodecoder_film(in_coord_features=3, in_latent_features=384, out_features=3,
                                 num_hidden_layers=15, hidden_features=384)  # scripts/train.py:230-236 style
    assert a._dims_tuple == b._dims_tuple == (3, 384, 384, 15, 3)
    b.load_state_dict(a.state_dict())
    a.disable_gradient()
    assert not any(p.requires_grad for p in a.parameters())


def test_unsupported_configs_raise():
    with pytest.raises(NotImplementedError):
        cb.SIRENAutodecoder_film(2, 8, 3, 1, 16, premap_mode="fourier")
    with pytest.raises(NotImplementedError):
        cb.SIRENAutodecoder_film(2, 8, 3, 1, 16, nonlinearity="relu")


def test_cpu_forward_fails_loudly():
    m = cb.SIRENAutodecoder_film(2, 8, 3, 1, 16).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 5, 2), torch.zeros(3, 1, 8))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 5, 3), torch.zeros(3, 1, 8))
    with pytest.raises(TypeError):
        m(torch.zeros(1, 5, 2, dtype=torch.float64), torch.zeros(3, 1, 8))


@pytest.mark.parametrize("cshape,lshape,T,P,stride", [
    ((1, 50, 2), (7, 1, 8), 7, 50, 0),         # pass_through_model_batch
    ((50, 2), (7, 1, 8), 7, 50, 0),            # trainer.infer, flat coords
    ((5, 6, 2), (7, 1, 1, 8), 7, 30, 0),       # CNF_inference.predict, grid coords
    ((7, 50, 2), (7, 1, 8), 7, 50, 100),       # training loop, per-frame coords
    ((50, 2), (8,), 1, 50, 0),                 # single latent vector
    ((4, 50, 2), (50, 8), 200, 1, 2),          # latents varying per point: every pair is its own frame
])
def test_canonicalize_matches_broadcast(cshape, lshape, T, P, stride):
    g = torch.Generator().manual_seed(0)
    c = torch.rand(cshape, generator=g)
    l = torch.rand(lshape, generator=g)
    cc, st, l2, T_, P_, lead = canonicalize(c, l)
    assert (T_, P_, st) == (T, P, stride)
    assert tuple(l2.shape) == (T, 8)
    # rebuild the broadcast operands from the canonical form and compare with torch broadcasting
    cfull = (cc[None].expand(T, P, 2) if st == 0 else cc).reshape(lead + (2,))
    lfull = l2[:, None, :].expand(T, P, 8).reshape(lead + (8,))
    assert torch.equal(cfull, c.expand(lead + (2,)))
    assert torch.equal(lfull, l.expand(lead + (8,)))


def test_canonicalize_keeps_latent_grad_path():
    l = torch.rand(3, 1, 8, requires_grad=True)
    _, _, l2, *_ = canonicalize(torch.rand(1, 5, 2), l)
    (g,) = torch.autograd.grad(l2.sum(), l)
    assert tuple(g.shape) == (3, 1, 8) and torch.all(g == 1)


def test_install_patches_named_modules():
    import types

    fake = types.ModuleType("fake.cnf.nf_networks")
    fake.SIRENAutodecoder_film = object
    fake.SIRENAutodecoder_film_extra_in = object
    assert cb.install(fake) == ["fake.cnf.nf_networks"]
    assert fake.SIRENAutodecoder_film is cb.SIRENAutodecoder_film
    assert fake.SIRENAutodecoder_film_extra_in is cb.SIRENAutodecoder_film_extra_in


class _OracleModel(torch.nn.Module):
    """Stand-in decode model for driver tests on CPU (checker only)."""

    def __init__(self, sd):
        super().__init__()
        self.sd = sd
        self.net1 = torch.nn.ModuleList([torch.nn.Linear(1, 1) for _ in range(O.dims_of(sd)[3] + 2)])
        self.net1[-1].weight = torch.nn.Parameter(sd[f"net1.{len(self.net1) - 1}.weight"].clone())

    def forward(self, coords, latents):
        return O.forward(self.sd, coords, latents)


class _Affine:
    def __init__(self, a, b):
        self.a, self.b = a, b

    def normalize(self, x):
        return x * self.a + self.b

    def denormalize(self, y):
        return (y - self.b) / self.a


def test_decode_drivers_equal_reference_loop():
    sd = O.init_params(2, 16, 3, 2, 32, seed=0)
    model = _OracleModel(sd)
    coords, lat = O.synthetic_inputs(2, 16, 11, 37)
    xn, yn = _Affine(0.5, 0.1), _Affine(2.0, -0.3)
    want = torch.cat([yn.denormalize(O.forward(sd, xn.normalize(coords[None]), lat[i:i + 4, None]))
                      for i in range(0, 11, 4)])  # the reference's loop, batch_size 4
    got = cb.pass_through_model_batch(coords, lat, model, xn, yn, 4, "cpu")
    assert torch.allclose(got, want, atol=1e-6)
    got2 = cb.decoder(coords, lat, model, xn, yn, 4, "cpu")
    assert torch.allclose(got2, want, atol=1e-6) and got2.device.type == "cpu"


class _RefNorm:
    """The reference's Normalizer_ts arithmetic (cnf/utils/normalize.py:100-120) for the folding test."""

    def __init__(self, method, params):
        self.method, self.params = method, params

    def normalize(self, x):
        p0, p1 = self.params
        return {"-11": lambda: (x - p1) / (p0 - p1) * 2 - 1, "01": lambda: (x - p1) / (p0 - p1),
                "ms": lambda: (x - p0) / p1, "none": lambda: x}[self.method]()

    def denormalize(self, y):
        p0, p1 = self.params
        return {"-11": lambda: (y + 1) / 2 * (p0 - p1) + p1, "01": lambda: y * (p0 - p1) + p1,
                "ms": lambda: y * p1 + p0, "none": lambda: y}[self.method]()


@pytest.mark.parametrize("xm,ym", [("-11", "-11"), ("01", "ms"), ("ms", "none"), ("none", "01")])
def test_fold_normalizers_matches_explicit_affines(xm, ym):
    """Folded parameters, evaluated by the oracle in fp64, equal normalise -> decode -> denormalise."""
    torch.manual_seed(0)
    m = cb.SIRENAutodecoder_film(2, 16, 3, 2, 32).double()
    g = torch.Generator().manual_seed(4)
    xn = _RefNorm(xm, (torch.tensor([2.0, 3.5], dtype=torch.float64), torch.tensor([-1.0, 0.5], dtype=torch.float64)))
    yn = _RefNorm(ym, (torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64), torch.tensor([0.5, 0.25, 2.0], dtype=torch.float64)))
    a_z = torch.rand(16, generator=g, dtype=torch.float64) + 0.5
    c_z = torch.randn(16, generator=g, dtype=torch.float64) * 0.1
    coords = torch.rand(1, 40, 2, generator=g, dtype=torch.float64) * 3 - 1
    lat = torch.randn(5, 1, 16, generator=g, dtype=torch.float64) * 0.2
    want = yn.denormalize(O.forward(m.state_dict(), xn.normalize(coords), a_z * lat + c_z))
    folded = cb.fold_normalizers(m, xn, yn, latent_affine=(a_z, c_z))
    got = O.forward(folded.state_dict(), coords, lat)
    assert torch.allclose(got, want, rtol=1e-10, atol=1e-10)
    assert not torch.equal(folded.net1[0].weight, m.net1[0].weight) or xm == "none"


@pytest.mark.parametrize("T,P,cout,bs", [(1024, 65536, 3, 16), (7, 300, 3, 16), (1, 1, 2, 1), (3, 65536, 4, 1),
                                         (2000, 1000, 3, 64), (16, 16384, 3, 4), (513, 10, 3, 256)])
def test_decoder_chunk_plan_covers_all_frames_and_tapers(T, P, cout, bs):
    from confild_b200.inference_function import _chunk_plan, _frames_per_chunk
    plan = _chunk_plan(T, P, cout, bs)
    step = _frames_per_chunk(T, P, cout, bs)
    assert sum(plan) == T and all(0 < n <= step for n in plan)
    # after the first (remainder) chunk sizes never grow, and the tail shrinks by at most 3x per chunk, so every
    # device->host copy hides under the next chunk's decode
    for a, b in zip(plan[1:], plan[2:]):
        assert b <= a and 3 * b >= a
