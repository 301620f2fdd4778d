"""CPU: the reference arm of bench.py (the reference's own decoder from oracle/_ref when it was built, else the oracle
port, on the host cores) prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cnf_decode_point_frames_per_s"
    assert d["unit"] == "point-frames/s" and d["higher_is_better"] is True and d["value"] > 0
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ConditionalNeuralField", "cnf", "nf_networks.bin"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("CNF decode case1")


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_measured_arm_generators_match_the_oracle():
    """bench.py's measured arm builds its weights and inputs without touching oracle/: same draws, bit for bit."""
    import importlib.util

    import torch

    import confild_b200 as cb
    from oracle import cnf_oracle as O

    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.CASE4_DIMS == O.CASE_SHAPES["case4"] and bench.DIMS == O.CASE_SHAPES["case1"]
    for a, b in zip(bench.synthetic_inputs(2, 128, 5, 77, latent_seed=4), O.synthetic_inputs(2, 128, 5, 77, latent_seed=4)):
        assert torch.equal(a, b)
    dims = (2, 16, 3, 2, 32)
    m = bench.seeded_model(cb, dims, "fp32")
    want = O.init_params(*dims, seed=0)
    got = m.state_dict()
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k


def test_compiled_reference_matches_the_port():
    """oracle/_ref (the reference's own modules, byte-compiled by oracle/build_ref.py) and the restated oracle give
    bit-identical weights and outputs; skipped where neither /root/reference nor an earlier build exists."""
    import sys as _sys

    import pytest
    import torch

    _sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    from oracle import cnf_oracle as O

    if not build_ref.build():
        pytest.skip("oracle/_ref not available (no /root/reference here and never built)")
    Ref = build_ref.load_reference_class()
    assert Ref.__module__ == "ConditionalNeuralField.cnf.nf_networks"
    torch.manual_seed(0)
    m = Ref(2, 128, 3, 10, 128).eval()
    sd = O.init_params(2, 128, 3, 10, 128, seed=0)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    c, l = O.synthetic_inputs(2, 128, 2, 61)
    assert torch.equal(m(c[None], l[:, None]), O.forward(sd, c[None], l[:, None]))
