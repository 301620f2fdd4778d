"""First-contact probe for the GPU box: runs each kernel family on small inputs, prints errors vs the
CPU oracle (checker only) and simple timings.  Not a test; used while bringing kernels up."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import confild_b200 as cb  # noqa: E402
from confild_b200 import _native  # noqa: E402
from oracle import cnf_oracle as O  # noqa: E402


def model(dims, sd, prec):
    cin, L, cout, nl, H = dims
    m = cb.SIRENAutodecoder_film(cin, L, cout, nl, H, precision=prec)
    m.load_state_dict(sd)
    return m.eval().cuda()


def run(case, T, P, precs, grad=True):
    dims = O.CASE_SHAPES[case] if isinstance(case, str) else case
    sd = O.init_params(*dims, seed=0)
    coords, lat = O.synthetic_inputs(dims[0], dims[1], T, P)
    want = O.forward(sd, coords[None], lat[:, None])
    gout = torch.randn(want.shape, generator=torch.Generator().manual_seed(7))
    gwant = O.grad_latents_from_gout(sd, coords[None], lat[:, None], gout) if grad else None
    for prec in precs:
        try:
            m = model(dims, sd, prec)
            with torch.no_grad():
                y = m(coords.cuda()[None], lat.cuda()[:, None])
            torch.cuda.synchronize()
            msg = f"{case} T={T} P={P} {prec}: fwd rel_l2 {O.rel_l2(y, want):.3e}"
            if grad:
                l = lat.cuda()[:, None].requires_grad_(True)
                y2 = m(coords.cuda()[None], l)
                (g,) = torch.autograd.grad(y2, l, grad_outputs=gout.cuda())
                torch.cuda.synchronize()
                msg += f" | fwd(stash) {O.rel_l2(y2, want):.3e} | dlat rel_l2 {O.rel_l2(g, gwant):.3e}"
            print(msg, flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{case} T={T} P={P} {prec}: FAILED {type(e).__name__}: {e}", flush=True)
            raise


def timeit(case, T, P, prec, iters=5):
    dims = O.CASE_SHAPES[case]
    sd = O.init_params(*dims, seed=0)
    coords, lat = O.synthetic_inputs(dims[0], dims[1], T, P)
    m = model(dims, sd, prec)
    c, l = coords.cuda()[None], lat.cuda()[:, None]
    with torch.no_grad():
        for _ in range(2):
            m(c, l)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            m(c, l)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"time {case} T={T} P={P} {prec}: {ms:.3f} ms  -> {T * P / ms / 1e6:.3f} G pf/s  launch={_native.query_launch(m._cdims(), m._precision_code(), T, P)}", flush=True)


def time_dps(case, T, P, prec, iters=5):
    dims = O.CASE_SHAPES[case]
    sd = O.init_params(*dims, seed=0)
    coords, lat = O.synthetic_inputs(dims[0], dims[1], T, P)
    m = model(dims, sd, prec)
    c = coords.cuda()[None]
    mask = torch.zeros(P, 1, device="cuda")
    mask[torch.randperm(P, device="cuda")[:min(P, 1000)]] = 1.0
    y_meas = torch.randn(T, P, dims[2], device="cuda") * 0.05

    def step():
        l = lat.cuda()[:, None].requires_grad_(True)
        y = m(c, l)
        loss = torch.linalg.norm((y_meas - y) * mask)
        (g,) = torch.autograd.grad(loss, l)
        return g

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # the same step captured in a CUDA graph (static latent buffer)
    static = lat.cuda()[:, None].clone().requires_grad_(True)

    def gstep():
        y = m(c, static)
        loss = torch.linalg.norm((y_meas - y) * mask)
        return torch.autograd.grad(loss, static)[0]

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            gstep()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        gstep()
    graph.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters * 4):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    gms = e0.elapsed_time(e1) / (iters * 4)
    print(f"time DPS fwd+bwd {case} T={T} P={P} {prec}: {ms:.3f} ms -> {T * P / ms / 1e6:.3f} G pf/s | CUDA graph replay {gms:.3f} ms", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    if what in ("all", "simt"):
        run((2, 32, 3, 2, 64), 3, 70, ["fp32"])
        run("case1", 3, 300, ["fp32"])
    if what in ("all", "tc"):
        run("case1", 1, 128, ["bf16x3"], grad=False)
        run("case1", 3, 300, ["bf16x3", "fp16"])
        run("case2", 2, 200, ["bf16x3", "fp16"])
        run("case4", 2, 200, ["bf16x3", "fp16"])
        run("case1", 16, 4099, ["bf16x3", "fp16"])
    if what in ("all", "time"):
        for prec in ("bf16x3", "fp16"):
            timeit("case1", 64, 65536, prec)
        for prec in ("bf16x3", "fp16"):
            timeit("case4", 16, 16384, prec)
        timeit("case1", 16, 16384, "fp32", iters=2)
    if what == "ragged":
        for P in (300, 200, 100, 40):
            timeit("case1", 256, P, "bf16x3")
        timeit("case4", 64, 300, "bf16x3")
    if what == "case2":
        for prec in ("bf16x3", "fp16"):
            timeit("case2", 32, 16384, prec)
    if what in ("all", "dps"):
        for prec in ("bf16x3", "fp16", "fp32"):
            time_dps("case1", 64, 16384, prec)
        for prec in ("bf16x3", "fp16"):
            time_dps("case4", 64, 16384, prec)
        time_dps("case4", 384, 1, "bf16x3")
        time_dps("case4", 384, 10, "bf16x3")
        time_dps("case4", 384, 100, "bf16x3")
        time_dps("case4", 384, 1000, "bf16x3")
