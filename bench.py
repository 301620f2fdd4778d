#!/usr/bin/env python
"""Headline benchmark: CNF decode throughput in field-points x frames per second.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                      # the reference's own decoder on the host cores

Workload (BASELINE.json configs[1]): case1 shapes (cin,L,H,nl,cout) = (2,128,128,10,3), 1024 frames x
65,536 query points per GPU, forward only, random-init weights (reference constructor order, seed 0),
synthetic coords ~ U(-1,1) and latents ~ N(0,0.1^2).  A step = one decode of all frames of the rank
(FiLM-shift GEMM + fused layer-chain kernel); for N > 1 every rank decodes its own 1024 frames (weak
scaling) and the decoded field is all-gathered inside the step (fused into the kernel's epilogue).
`extra.config3_case4_sharded` is BASELINE.json configs[2]: case4 shapes, 4,096 frames x 131,072 points
in total, frames sharded over the N ranks (strong scaling), gather inside the step.
Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CASE = "case1"
DIMS = (2, 128, 3, 10, 128)  # cin, L, cout, nl, H (oracle order)
CASE4_DIMS = (3, 384, 3, 15, 384)  # the 3-D recipe (BASELINE configs 3 and 4)
FRAMES, POINTS = 1024, 65536
CONFIG3_FRAMES, CONFIG3_POINTS = 4096, 131072
METRIC = "cnf_decode_point_frames_per_s"
UNIT = "point-frames/s"
#: chip-wide sin.approx (MUFU.SIN) throughput measured on this pool's B200 with scripts/microbench.cu
#: (profiles/r01_microbench_mma_ldtm_mufu.txt): 15.98 sin/clk/SM, 4.625 T sin/s at 1.965 GHz
MUFU_PEAK_SIN_PER_S = 4.625e12
#: DRAM bytes (read + write) of ONE launch of the headline kernel at the bench size, from an offline `ncu --set full`
#: capture of this command (not measured inside this run): {precision: (bytes, profile file)}
NCU_TRAFFIC_OFFLINE = {
    "bf16x3": (771521280, "profiles/r01_ncu_tc2_forward_case1_bf16x3_benchsize_final.txt"),
    "f16f8": (776403968, "profiles/r02_ncu_tc2_forward_case1_f16f8_benchsize.txt"),  # 16.25 MB read + 760.15 MB written
}


def flops_per_pf(cin, L, cout, nl, H):
    return 2 * (cin * H + nl * H * H + H * cout)


def sins_per_pf(cin, L, cout, nl, H):
    return (nl + 1) * H


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


def roofline_block(dims, T, P, kernel_ms, peaks, peak_src, precision, kernel):
    """Both candidate ceilings of SURVEY.md 8(d) -- algorithmic FLOPs against the measured sustained bf16 tensor
    peak, algorithmic sines against the measured MUFU peak -- and the rule that names the binding one:
    bound = argmax(F_alg / P_tensor, S_alg / P_mufu), i.e. the resource whose ideal time is longer."""
    fl = flops_per_pf(*dims) * T * P
    sn = sins_per_pf(*dims) * T * P
    ach_tf = fl / (kernel_ms * 1e-3) / 1e12
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    sin_rate = sn / (kernel_ms * 1e-3)
    t_tensor, t_mufu = fl / (peak_tf * 1e12), sn / MUFU_PEAK_SIN_PER_S
    binding = "tensor" if t_tensor >= t_mufu else "mufu"
    tensor_frac, mufu_frac = ach_tf / peak_tf, sin_rate / MUFU_PEAK_SIN_PER_S
    traffic, traffic_src = NCU_TRAFFIC_OFFLINE.get(precision, (None, None)) if (T, P) == (FRAMES, POINTS) and dims == DIMS \
        else (None, None)
    return {
        "bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tensor_frac,
        "traffic": traffic, "traffic_source": ("ncu --set full, offline capture, not measured in this run: " + traffic_src)
        if traffic_src else None,
        "kernel": kernel, "kernel_ms": kernel_ms, "peak_source": f"{peak_src} bf16_tflops_sustained",
        "algorithmic_flops_per_launch": fl, "algorithmic_sines_per_launch": sn,
        "tensor_frac": tensor_frac, "mufu_frac": mufu_frac, "binding": binding,
        "binding_frac": tensor_frac if binding == "tensor" else mufu_frac,
        "binding_rule": "argmax(F_alg/P_tensor, S_alg/P_mufu): ideal tensor time %.3f ms vs ideal MUFU time %.3f ms per "
                        "launch" % (t_tensor * 1e3, t_mufu * 1e3),
        "mufu": {"achieved_gsin_s": sin_rate / 1e9, "measured_peak_gsin_s": MUFU_PEAK_SIN_PER_S / 1e9, "frac": mufu_frac},
    }


# ------------------------------------------------------------------------------------------ CPU arm
def reference_decoder(dims):
    """(decode(coords (P,cin), latents (T,L)) -> (T,P,cout), kind, description) on the host.

    kind "reference": the reference's OWN SIRENAutodecoder_film (byte-compiled from /root/reference by
    oracle/build_ref.py into oracle/_ref, which travels to the GPU box); kind "port": the restated oracle, when
    oracle/_ref was never built.  Same seed-0 constructor weights either way (bit-identical, tests/test_oracle.py)."""
    cin, L, cout, nl, H = dims
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    why = "oracle/_ref not built"
    try:
        import build_ref  # the one other place bench.py executes oracle/

        Ref = build_ref.load_reference_class()
    except Exception as e:  # noqa: BLE001
        Ref, why = None, f"oracle/_ref failed to load: {type(e).__name__}: {e}"[:120]
    if Ref is not None:
        torch.manual_seed(0)
        model = Ref(cin, L, cout, nl, H).eval()

        def decode(coords, lat):
            with torch.no_grad():
                return model(coords[None], lat[:, None])

        return decode, "reference", "reference SIRENAutodecoder_film.forward (oracle/_ref bytecode of cnf/nf_networks.py)"
    from oracle import cnf_oracle as O

    sd = O.init_params(*dims, seed=0)

    def decode_port(coords, lat):
        with torch.no_grad():
            return O.forward(sd, coords[None], lat[:, None])

    return decode_port, "port", f"oracle/cnf_oracle.py restatement ({why})"


def synthetic_inputs(cin, L, T, P, sigma=0.1, coord_seed=1, latent_seed=2):
    """coords ~ U(-1,1) (P,cin), generator seed 1; latents ~ N(0, sigma^2) (T,L), generator seed 2 (SURVEY.md 8(d) 'Inputs').
    Same draws as the oracle's generator (tests/test_bench_cpu.py checks that), kept here so that the measured arm does not
    touch oracle/."""
    gc = torch.Generator().manual_seed(coord_seed)
    gl = torch.Generator().manual_seed(latent_seed)
    return torch.rand(P, cin, generator=gc) * 2 - 1, torch.randn(T, L, generator=gl) * sigma


def seeded_model(cb, dims, precision):
    """Random-init weights of the named architecture: the module's own constructor under torch.manual_seed(0) draws them
    in the reference constructor's order (bit-identical to the reference, tests/test_host.py)."""
    cin, L, cout, nl, H = dims
    torch.manual_seed(0)
    return cb.SIRENAutodecoder_film(cin, L, cout, nl, H, precision=precision)


def run_cpu_baseline(sample_frames=16, reps=3):
    """The reference decode on all host cores, bounded sample of the headline workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    decode, kind, what = reference_decoder(DIMS)
    coords, lat = synthetic_inputs(DIMS[0], DIMS[1], sample_frames, POINTS)
    decode(coords, lat[:1])  # warm-up
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        decode(coords, lat)
        best = min(best, time.perf_counter() - t0)
    return {"value": sample_frames * POINTS / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{CASE} shapes, {sample_frames} frames x {POINTS} points, fp32, {what}, torch {torch.__version__} "
                      f"CPU, best of {reps} after warm-up ({best:.2f} s)"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    decode, kind, what = reference_decoder(DIMS)
    coords, lat_all = synthetic_inputs(DIMS[0], DIMS[1], 16, POINTS)
    t0 = time.perf_counter()
    decode(coords, lat_all[:1])
    t1 = time.perf_counter() - t0
    budget = 150.0 / max(1, args.steps + args.warmup)
    sample = 16
    while sample > 1 and sample * t1 > budget:
        sample //= 2
    lat = lat_all[:sample]
    for _ in range(args.warmup):
        decode(coords, lat)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        decode(coords, lat)
    dt = time.perf_counter() - t0
    value = args.steps * sample * POINTS / dt
    sample_txt = f"{CASE} shapes, {sample} frames x {POINTS} points per step, fp32, all host threads, {what}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CNF decode {CASE} 2D shapes, forward only (CPU: bounded sample of the 1024-frame job)",
                   "frames_per_step": sample, "points": POINTS, "dims": dict(zip(("cin", "L", "cout", "nl", "H"), DIMS))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, reasons, power = [], set(), []
        smax = None
        for ln in self.tmp.read().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            busy = sorted(sm)[len(sm) // 2:]  # upper half = samples under load
            out.update(sm_mhz=busy[len(busy) // 2], sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power) if power else None)
        return out


class Affine11:
    """'-11' normaliser of the reference (cnf/utils/normalize.py:100-120) with fixed (max, min)."""
    method = "-11"

    def __init__(self, hi, lo):
        self.params = (hi, lo)

    def normalize(self, x):
        hi, lo = (p.to(x.device) for p in self.params)
        return (x - lo) / (hi - lo) * 2 - 1

    def denormalize(self, y):
        hi, lo = (p.to(y.device) for p in self.params)
        return (y + 1) / 2 * (hi - lo) + lo


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def eager_decode(params, coords, lat, w0=30.0):
    """The reference's forward as plain PyTorch eager ops on the GPU (nf_networks.py:491-494: matmul, in-place bias
    add, broadcast add, mul, sin per layer) -- the 'PyTorch-eager on the same B200' comparator of SURVEY.md 2.1."""
    net1, net2 = params
    x = coords
    for i in range(len(net2)):
        w, b = net1[i]
        h = torch.matmul(x, w.t())
        h += b
        x = torch.sin(w0 * (h + torch.matmul(lat, net2[i].t())))
    w, b = net1[-1]
    return torch.matmul(x, w.t()) + b


def measure_eager(model, dev):
    nl = model._dims_tuple[3]
    net1 = [(model.net1[i].weight.detach(), model.net1[i].bias.detach()) for i in range(nl + 2)]
    net2 = [model.net2[i].weight.detach() for i in range(nl + 1)]
    T, P = 64, POINTS
    c, l = synthetic_inputs(DIMS[0], DIMS[1], T, P)
    c, l = c.to(dev)[None], l.to(dev)[:, None]
    out = {"workload": f"{CASE} shapes, {T} frames x {P} points, forward only, torch {torch.__version__} eager on the same GPU"}
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                ms = timed(lambda: eager_decode((net1, net2), c, l), 3)
            out[name] = {"value": T * P / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


def measure_dps(cb, model, dims, case, dev, args, iters):
    """BASELINE config 4: 64 latents x 16,384 points, 1,000 random sensors.  Three ways, never folded into the headline:
    (a) the reference's formulation through the drop-in module and PyTorch autograd (dense dL/dy);
    (b) the fused measurement norm (cnf_forward_loss, no element-wise PyTorch pass), still dense over all P;
    (c) sensor-compacted: decode + backward on the 1,000 sensor rows only (what Case3/4Operator do by passing sensor coords)."""
    cin, L, cout, nl, H = dims
    Td, Pd, S = 64, 16384, 1000
    cd, ld = synthetic_inputs(cin, L, Td, Pd)
    cd, ld = cd.to(dev), ld.to(dev)
    mask = torch.zeros(Pd, device=dev)
    mask[torch.randperm(Pd, device=dev)[:S]] = 1.0
    y_meas = torch.randn(Td, Pd, cout, device=dev) * 0.05
    ym_masked = y_meas * mask[None, :, None]

    def autograd_step():
        l = ld[:, None].detach().requires_grad_(True)
        y = model(cd[None], l)
        loss = torch.linalg.norm((y_meas - y) * mask[None, :, None])  # condition_methods.py:30-31
        torch.autograd.grad(loss, l)                                   # condition_methods.py:32

    def fused_step():
        l = ld[:, None].detach().requires_grad_(True)
        norm = cb.measurement_norm(model, cd[None], l, ym_masked, mask=mask, zero_row_skip=False)
        torch.autograd.grad(norm, l)

    def fused_skip_step():
        l = ld[:, None].detach().requires_grad_(True)
        norm = cb.measurement_norm(model, cd[None], l, ym_masked, mask=mask, skip_masked_decode=False)
        torch.autograd.grad(norm, l)

    def fused_default_step():  # measurement_norm's default: rows with a zero mask weight are neither decoded nor stashed
        l = ld[:, None].detach().requires_grad_(True)
        norm = cb.measurement_norm(model, cd[None], l, y_meas, mask=mask)
        torch.autograd.grad(norm, l)

    graphed = cb.GraphedMeasurementNorm(model, cd[None], ld[:, None], y_meas, mask=mask)

    def fused_graph_step():  # the same step captured once in a CUDA graph (GraphedMeasurementNorm) and replayed
        l = ld[:, None].detach().requires_grad_(True)
        torch.autograd.grad(graphed(l), l)

    cs, _, ys = cb.sensor_rows(cd, mask, y_meas)

    def compact_step():
        l = ld[:, None].detach().requires_grad_(True)
        norm = cb.measurement_norm(model, cs[None], l, ys)
        torch.autograd.grad(norm, l)

    res = {"workload": f"{case} shapes, {Td} latents x {Pd} points, {S} random sensors, forward(+cos stash) + loss + "
                       "backward to dL/dlatent (BASELINE config 4)", "precision": model.precision,
           "backward_precision": "fp32" if model.resolved_precision == "fp32" else
           "bf16x3 (the backward always runs the bf16 hi/lo split on the fp16 cos stash, whatever the forward mode)"}
    for name, fn, rows in (("autograd_dense", autograd_step, Td * Pd), ("fused_loss_dense", fused_step, Td * Pd),
                           ("fused_loss_dense_zero_row_skip", fused_skip_step, Td * Pd),
                           ("fused_loss_masked_rows_not_decoded", fused_default_step, Td * S),
                           ("fused_loss_masked_rows_not_decoded_cuda_graph", fused_graph_step, Td * S),
                           ("sensor_compacted", compact_step, Td * S)):
        ms = timed(fn, iters)
        res[name] = {"ms_per_step": ms, "point_frames_per_s": rows / (ms * 1e-3), "rows_per_step": rows}
    res["value"] = res["fused_loss_dense"]["point_frames_per_s"]
    res["ms_per_step"] = res["fused_loss_dense"]["ms_per_step"]
    res["unit"] = UNIT
    res["note"] = ("value = fused_loss_dense: every point decoded WITH its backward stash and visited by the backward (the "
                   "graded dense quantity); fused_loss_dense_zero_row_skip also decodes and scores every point but stashes "
                   "/ back-propagates only the rows whose mask weight is non-zero (exact gradient, P/#sensors less stash "
                   "traffic); fused_loss_masked_rows_not_decoded = measurement_norm's default on the same full-grid inputs "
                   "(unmasked measurement, per-point mask): a row with a zero weight has residual = measurement whatever the "
                   "decoder returns, so only the sensor rows are decoded / stashed / back-propagated and the other rows' "
                   "measurement energy is added to the sum of squares -- same norm and gradient, rows_per_step counts the "
                   "decoded rows; sensor_compacted is the caller-side version of the same thing (SURVEY.md 8d: reported "
                   "separately)")
    return res


def measure_extra(model, coords, lat, dev, args):
    """Reported next to the headline (never folded into it): the other precisions on the same workload, the DPS step of
    BASELINE config 4 for case1 and case4 shapes, the notebook's literal DPS shape, a bounded case4 decode and the
    PyTorch-eager comparator on the same GPU."""
    import confild_b200 as cb

    cin, L, cout, nl, H = DIMS
    T, P = args.frames, args.points
    out = {}
    for prec in ("bf16x3", "f16f8", "fp16"):
        if prec == args.precision:
            continue
        other = cb.SIRENAutodecoder_film(cin, L, cout, nl, H, precision=prec)
        other.load_state_dict(model.state_dict())
        other = other.eval().to(dev)
        with torch.no_grad():
            ms = timed(lambda: other(coords, lat), 3)
        out[f"precision_{prec}"] = {"value": T * P / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms}
        del other
    model.disable_gradient()
    out["dps_fwd_bwd"] = measure_dps(cb, model, DIMS, CASE, dev, args, 5)
    out["eager_b200"] = measure_eager(model, dev)

    # ---- case4 (3-D recipe: 15 hidden layers of width 384), the shapes of BASELINE configs 3 and 4
    dims4 = CASE4_DIMS
    m4 = seeded_model(cb, dims4, args.precision).eval().to(dev)
    m4.disable_gradient()
    out["dps_fwd_bwd_case4"] = measure_dps(cb, m4, dims4, "case4", dev, args, 3)
    # the notebook's literal shape: 384 frames x 10 sensor points (ipynb :194-200, :304-306)
    Tn, Pn = 384, 10
    cn, ln = synthetic_inputs(dims4[0], dims4[1], Tn, Pn)
    cn, ln = cn.to(dev), ln.to(dev)
    ymn = torch.randn(Tn, Pn, dims4[2], device=dev) * 0.05

    def nb_autograd():
        l = ln[:, None].detach().requires_grad_(True)
        torch.autograd.grad(torch.linalg.norm(ymn - m4(cn[None], l)), l)

    def nb_fused():
        l = ln[:, None].detach().requires_grad_(True)
        torch.autograd.grad(cb.measurement_norm(m4, cn[None], l, ymn), l)

    nb_graphed_fn = cb.GraphedMeasurementNorm(m4, cn[None], ln[:, None], ymn)

    def nb_graphed():
        l = ln[:, None].detach().requires_grad_(True)
        torch.autograd.grad(nb_graphed_fn(l), l)

    out["dps_notebook_shape_case4"] = {
        "workload": f"case4 shapes, {Tn} frames x {Pn} sensor points (the notebook's literal DPS shape)",
        "autograd_ms": timed(nb_autograd, 10), "fused_loss_ms": timed(nb_fused, 10),
        "fused_loss_cuda_graph_ms": timed(nb_graphed, 10)}
    return out


def measure_config3(cb, dev, world, rank, args):
    """BASELINE.json configs[2]: case4 shapes, 4,096 frames x 131,072 points in total, frames sharded over the ranks
    (strong scaling), decoded field gathered to every rank inside the step (fused into the kernel's epilogue for N > 1).
    Device-timed, max over ranks."""
    import torch.distributed as dist

    dims4 = CASE4_DIMS
    T_total, P = args.config3_frames, CONFIG3_POINTS
    if T_total % world:
        return {"skipped": f"{T_total} frames do not split evenly over {world} ranks"}
    T = T_total // world
    m4 = seeded_model(cb, dims4, args.precision).eval().to(dev)
    ev = []
    m4._timing = ev
    c4, l4 = synthetic_inputs(dims4[0], dims4[1], T_total, P)
    c4, l4 = c4.to(dev)[None], l4[rank * T:(rank + 1) * T].to(dev)[:, None]
    fused = None
    mode = "single GPU, no gather"
    if world > 1:
        fused = cb.FusedGatherDecoder(m4, T, P)
        mode = "all-gather fused into the decode kernel (peer stores over NVLink, double-buffered symmetric memory)"

    def step():
        with torch.no_grad():
            return fused(c4, l4) if fused is not None else m4(c4, l4)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    step()
    barrier()
    ev.clear()
    steps = args.config3_steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms_local = e0.elapsed_time(e1) / steps
    kern_ms = sum(a.elapsed_time(b) for a, b in ev) / max(1, len(ev))
    t = torch.tensor([ms_local, kern_ms], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        allms = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allms, t)
        per_rank = [float(x[0]) for x in allms]
    else:
        per_rank = [ms_local]
    ms = float(tmax[0])
    peaks, peak_src = load_peaks()
    res = {"value": T_total * P / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "scaling": "strong",
           "frames_total": T_total, "frames_per_gpu": T, "points": P, "precision": args.precision,
           "dims": dict(zip(("cin", "L", "cout", "nl", "H"), dims4)), "gather": mode, "per_rank_ms": per_rank,
           "workload": "BASELINE.json configs[2]: case4 3-D shapes, 4096 frames x 131072 points, frame-sharded",
           "roofline": roofline_block(dims4, T, P, float(tmax[1]), peaks, peak_src, args.precision, "tc_forward_kernel<384>")}
    m4._timing = None
    del fused, m4
    torch.cuda.empty_cache()
    return res


def measure_config5(cb, model, dev, world, rank, args):
    """BASELINE.json configs[4]: full unconditional generation -- 1,000-step DDPM sampling of the latent image with the
    case1 U-Net (training_recipes/case1.yml: 128 x 128, 128 channels, 2 res blocks, attention at 32/16/8; random-init
    weights: no checkpoint offline) followed by the CNF decode of every generated frame.  16 samples over 8 GPUs in the
    reference recipe = 2 samples per rank, kept per rank at every N (weak scaling).  Sampler and decode are timed
    separately (the decode is the graded part); the sampler's per-step time is also given for plain fp32 eager PyTorch
    (what the reference runs), measured over a few steps."""
    import torch.distributed as dist

    per_rank, Tl, Ll = 2, 128, 128
    steps = args.config5_steps
    torch.manual_seed(1234 + rank)
    unet = cb.LatentUNet(image_size=128, num_channels=128, num_res_blocks=2, num_heads=4, num_head_channels=64,
                         attention_resolutions="32,16,8").eval().to(dev)
    coords_h, _ = synthetic_inputs(DIMS[0], DIMS[1], 1, POINTS)
    xn = Affine11(torch.tensor([1.0, 1.0]), torch.tensor([-1.0, -1.0]))
    yn = Affine11(torch.tensor([2.0, 1.5, 1.0]), torch.tensor([-2.0, -1.5, -1.0]))
    hi, lo = torch.full((Ll,), 0.3), torch.full((Ll,), -0.3)
    out_h = torch.empty((per_rank * Tl, POINTS, DIMS[2]), dtype=torch.float32, pin_memory=True)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    cb.sample_latents(unet, (per_rank, 1, Tl, Ll), steps=3, device=dev)  # warm-up (autotune, graph capture path)
    sync()
    t0 = time.perf_counter()
    z = cb.sample_latents(unet, (per_rank, 1, Tl, Ll), steps=steps, device=dev)
    sync()
    t_sample = time.perf_counter() - t0
    lat = ((z[:, 0] + 1) * (hi.to(dev) - lo.to(dev)) / 2 + lo.to(dev)).reshape(per_rank * Tl, Ll)
    cb.decoder(coords_h, lat, model, xn, yn, 16, dev, out=out_h)
    sync()
    t0 = time.perf_counter()
    cb.decoder(coords_h, lat, model, xn, yn, 16, dev, out=out_h)
    sync()
    t_decode = time.perf_counter() - t0
    # the reference's way of running the sampler: fp32 eager, one Python-driven step at a time
    n_e = 10
    cb.sample_latents(unet, (per_rank, 1, Tl, Ll), steps=2, device=dev, autocast_dtype=None, use_cuda_graph=False)
    sync()
    t0 = time.perf_counter()
    cb.sample_latents(unet, (per_rank, 1, Tl, Ll), steps=n_e, device=dev, autocast_dtype=None, use_cuda_graph=False)
    sync()
    t_eager_step = (time.perf_counter() - t0) / n_e
    tt = torch.tensor([t_sample, t_decode, t_eager_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_sample, t_decode, t_eager_step = (float(v) for v in tt)
    frames = world * per_rank * Tl
    del unet
    torch.cuda.empty_cache()
    return {"workload": f"BASELINE.json configs[4]: {world * per_rank} samples ({per_rank} per GPU) x {steps}-step DDPM of a "
                        f"(1,{Tl},{Ll}) latent image with the case1 U-Net (92 M parameters, random init), then CNF decode of "
                        f"{frames} frames x {POINTS} points to pinned host memory",
            "sampler_s": t_sample, "sampler_ms_per_step": t_sample / steps * 1e3,
            "sampler": "LatentUNet.forward_inference: bf16 channels-last end to end, GroupNorm(+embedding add)+SiLU as the "
                       "library's cnf_group_norm_nhwc_bf16 kernels, fused attention, one step captured in a CUDA graph",
            "sampler_eager_fp32_ms_per_step": t_eager_step * 1e3,
            "sampler_speedup_vs_eager_fp32": t_eager_step / (t_sample / steps),
            "decode_s": t_decode, "decode_point_frames_per_s": frames * POINTS / t_decode, "unit": UNIT,
            "total_s": t_sample + t_decode, "samples": world * per_rank, "steps": steps}


def main_ours(args):
    import torch.distributed as dist

    import confild_b200 as cb
    from confild_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base = run_cpu_baseline()  # before any CUDA work, same process
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)  # NCCL's own logging is left to the environment (stderr)

    cin, L, cout, nl, H = DIMS
    T, P = args.frames, args.points
    model = seeded_model(cb, DIMS, args.precision).eval().to(dev)
    requested = args.precision
    args.precision = model.resolved_precision  # "auto" -> the mode that actually runs; printed in dtype / config
    coords_h, lat_h = synthetic_inputs(cin, L, T, P, latent_seed=2 + rank)
    coords_h, lat_h = coords_h.pin_memory(), lat_h.pin_memory()
    coords, lat = coords_h.to(dev)[None], lat_h.to(dev)[:, None]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    kernel_events = []
    model._timing = kernel_events
    gather_mode = "none"
    fused = gathered = None
    if world > 1:
        try:  # all-gather fused into the decode kernel (peer stores over NVLink from the epilogue)
            if args.gather == "nccl":
                raise RuntimeError("NCCL all-gather requested")
            fused = cb.FusedGatherDecoder(model, T, P)
            gather_mode = ("fused into the decode kernel: epilogue stores to all ranks' symmetric-memory buffers over "
                           "NVLink (double-buffered)")
        except Exception as e:  # noqa: BLE001 - symmetric memory unavailable: decode + NCCL all_gather_into_tensor
            fused = None
            gathered = torch.empty((world * T, P, cout), dtype=torch.float32, device=dev)
            gather_mode = f"decode + NCCL all_gather_into_tensor ({type(e).__name__}: {e})"[:200]

    def step():
        with torch.no_grad():
            if fused is not None:
                return fused(coords, lat)
            y = model(coords, lat)
            if world > 1:
                dist.all_gather_into_tensor(gathered, y)
        return y

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    kernel_events.clear()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()  # evict L2 between timed steps (outside the per-step events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = sum(a.elapsed_time(b) for a, b in evs)
    kern_ms = [a.elapsed_time(b) for a, b in kernel_events]
    model._timing = None
    t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t.item())
    value = world * T * P * args.steps / (step_ms * 1e-3)
    del fused, gathered
    torch.cuda.empty_cache()

    # ---- end to end through the reference-facing driver: host buffers in, host field out
    xn = Affine11(torch.tensor([1.0, 1.0]), torch.tensor([-1.0, -1.0]))
    yn = Affine11(torch.tensor([2.0, 1.5, 1.0]), torch.tensor([-2.0, -1.5, -1.0]))
    out_h = torch.empty((T, P, cout), dtype=torch.float32, pin_memory=True)
    e2e_steps = max(2, min(args.steps, 5))
    cb.decoder(coords_h, lat_h, model, xn, yn, 16, dev, out=out_h)  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        cb.decoder(coords_h, lat_h, model, xn, yn, 16, dev, out=out_h)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * T * P * e2e_steps / float(t.item())
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(coords_h.numel() * 4 + lat_h.numel() * 4),
           "d2h_bytes_per_step": int(out_h.numel() * 4), "steps": e2e_steps,
           "api": "confild_b200.decoder(coords, latents, model, x_normalizer, y_normalizer, 16, device) "
                  "with pinned host buffers (mirror of cnf/inference_function.py:51-76)"}
    # the wall the e2e number runs into: the same bytes device->host with NO decode, all ranks at once (max over ranks)
    y_dev = torch.empty((T, P, cout), dtype=torch.float32, device=dev)
    out_h.copy_(y_dev, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        out_h.copy_(y_dev, non_blocking=True)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    d2h_gbs = 3 * out_h.numel() * 4 / float(t.item()) / 1e9
    e2e["d2h_only"] = {"gb_per_s_per_gpu": d2h_gbs, "gb_per_s_all_ranks": d2h_gbs * world,
                       "point_frames_per_s_ceiling": world * T * P * 3 / float(t.item()),
                       "note": "copy of the decoded field to pinned host memory alone, all ranks concurrently: the "
                               "host-side ceiling of the e2e metric"}
    del y_dev

    extra = {}
    if not args.no_extra:
        if world == 1:
            extra = measure_extra(model, coords, lat, dev, args)
        del coords, lat, flush
        torch.cuda.empty_cache()
        extra["config3_case4_sharded"] = measure_config3(cb, dev, world, rank, args)
        if args.config5_steps > 0:
            extra["config5_generation"] = measure_config5(cb, model, dev, world, rank, args)

    if rank == 0:
        peaks, peak_src = load_peaks()
        kavg_ms = sum(kern_ms) / max(1, len(kern_ms))
        launch = _native.query_launch(model._cdims(), model._precision_code(), T, P)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"CNF decode {CASE} 2D shapes, {T} frames x {P} points per GPU, forward only "
                                   "(BASELINE.json configs[1])",
                       "dims": dict(zip(("cin", "L", "cout", "nl", "H"), DIMS)), "frames_per_gpu": T, "points": P,
                       "precision": cb.PRECISION_NOTES[args.precision],
                       "precision_requested": requested,
                       "precision_policy": cb.PRECISION_NOTES["auto"] + "; measured forward rel-L2 vs the reference: "
                                           "f16f8 2.6e-5 (case1) .. 1.1e-4 (case3), bf16x3 5.6e-6 .. 2.4e-5, fp16 3.9e-4 .. "
                                           "1.8e-3 (tests/test_gpu_parity.py, DESIGN.md section 3)",
                       "parallelism": f"frames sharded over {world} GPU(s)" + (f"; all-gather of the field inside the step, {gather_mode}" if world > 1 else ""),
                       "l2": "256 MiB memset between timed steps; each step also writes %.0f MB of output" % (T * P * cout * 4 / 1e6)},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": 2 * args.steps,
            "roofline": roofline_block(DIMS, T, P, kavg_ms, peaks, peak_src, args.precision,
                                       "tc2_forward_kernel" if args.precision != "fp32" else "simt_forward_kernel"),
            "launch": dict(zip(("sms", "ctas", "threads", "smem_bytes", "ctas_per_sm", "tmem_cols", "tile_points"), launch)),
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if extra:
            line["extra"] = extra
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        # the single JSON line is the LAST thing written to stdout: NCCL's own logging (NCCL_DEBUG is left to the
        # environment so that the driver can read the communicator lines) comes before it
        sys.stdout.flush()
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CONFILD_PRECISION", "auto"),
                    choices=["auto", "f16f8", "bf16x3", "fp16", "fp32"],
                    help="operand precision; 'auto' = the module's policy (f16f8 on the tensor-core shapes)")
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--points", type=int, default=POINTS)
    ap.add_argument("--config3-frames", type=int, default=CONFIG3_FRAMES, help="total frames of extra.config3_case4_sharded")
    ap.add_argument("--config3-steps", type=int, default=3)
    ap.add_argument("--config5-steps", type=int, default=1000, help="DDPM steps of extra.config5_generation (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the side measurements (other precisions, DPS, case4, config 3)")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="N>1: how the decoded field is all-gathered")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
