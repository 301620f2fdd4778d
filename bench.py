#!/usr/bin/env python
"""Headline benchmark: CNF decode throughput in field-points x frames per second.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                      # the reference algorithm on the host cores

Workload (BASELINE.json configs[1]): case1 shapes (cin,L,H,nl,cout) = (2,128,128,10,3), 1024 frames x
65,536 query points per GPU, forward only, random-init weights (reference constructor order, seed 0),
synthetic coords ~ U(-1,1) and latents ~ N(0,0.1^2).  A step = one decode of all frames of the rank
(FiLM-shift GEMM + fused layer-chain kernel); for N > 1 every rank decodes its own 1024 frames (weak
scaling) and the decoded field is all-gathered over NCCL inside the step.
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
FRAMES, POINTS = 1024, 65536
METRIC = "cnf_decode_point_frames_per_s"
UNIT = "point-frames/s"
#: chip-wide sin.approx (MUFU.SIN) throughput measured on this pool's B200 with scripts/microbench.cu
#: (profiles/r01_microbench_mma_ldtm_mufu.txt): 15.98 sin/clk/SM, 4.625 T sin/s at 1.965 GHz
MUFU_PEAK_SIN_PER_S = 4.625e12
#: DRAM bytes of one tc2_forward_kernel launch at the bench size from `ncu --set full` (profiles/): read + write
NCU_TRAFFIC_BYTES_PER_LAUNCH = 771521280  # 14.64 MB read + 756.88 MB written (profiles/r01_ncu_tc2_forward_case1_bf16x3_benchsize_final.txt)


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


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(sd, coords, lat):
    from oracle import cnf_oracle as O  # the one place bench.py executes the oracle

    with torch.no_grad():
        return O.forward(sd, coords[None], lat[:, None])


CASE4_DIMS = (3, 384, 3, 15, 384)  # the 3-D recipe (BASELINE configs 3 and 4)


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
    """Oracle port of the reference decode (identical op order) on all host cores, bounded sample."""
    from oracle import cnf_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.init_params(*DIMS, seed=0)
    coords, lat = O.synthetic_inputs(DIMS[0], DIMS[1], sample_frames, POINTS)
    cpu_reference_step(sd, coords, lat[:1])  # warm-up
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_reference_step(sd, coords, lat)
        best = min(best, time.perf_counter() - t0)
    return {"value": sample_frames * POINTS / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{CASE} shapes, {sample_frames} frames x {POINTS} points, fp32, torch {torch.__version__} CPU, "
                      f"best of {reps} after warm-up ({best:.2f} s)"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cnf_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.init_params(*DIMS, seed=0)
    coords, lat_all = O.synthetic_inputs(DIMS[0], DIMS[1], 16, POINTS)
    t0 = time.perf_counter()
    cpu_reference_step(sd, coords, lat_all[:1])
    t1 = time.perf_counter() - t0
    budget = 150.0 / max(1, args.steps + args.warmup)
    sample = 16
    while sample > 1 and sample * t1 > budget:
        sample //= 2
    lat = lat_all[:sample]
    for _ in range(args.warmup):
        cpu_reference_step(sd, coords, lat)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sd, coords, lat)
    dt = time.perf_counter() - t0
    value = args.steps * sample * POINTS / dt
    sample_txt = f"{CASE} shapes, {sample} frames x {POINTS} points per step, fp32, all host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CNF decode {CASE} 2D shapes, forward only (CPU: bounded sample of the 1024-frame job)",
                   "frames_per_step": sample, "points": POINTS, "dims": dict(zip(("cin", "L", "cout", "nl", "H"), DIMS))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
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

    def __init__(self, hi, lo):
        self.params = (hi, lo)

    def normalize(self, x):
        hi, lo = (p.to(x.device) for p in self.params)
        return (x - lo) / (hi - lo) * 2 - 1

    def denormalize(self, y):
        hi, lo = (p.to(y.device) for p in self.params)
        return (y + 1) / 2 * (hi - lo) + lo


def measure_extra(model, coords, lat, dev, args):
    """Reported next to the headline (never folded into it): the single-pass fp16 fast mode on the same workload, the DPS
    step of BASELINE config 4 (forward with stash + backward to the latents, 64 frames x 16,384 points) for case1 and case4
    shapes, and a case4 decode (config 3 shapes)."""
    import confild_b200 as cb

    cin, L, cout, nl, H = DIMS
    T, P = args.frames, args.points
    out = {}

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

    if args.precision != "fp16":
        fast = cb.SIRENAutodecoder_film(cin, L, cout, nl, H, precision="fp16")
        fast.load_state_dict(model.state_dict())
        fast = fast.eval().to(dev)
        with torch.no_grad():
            ms = timed(lambda: fast(coords, lat), 3)
        out["fast_mode_fp16"] = {"value": T * P / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                 "note": "single fp16 MMA per product; forward rel-L2 vs reference 3.8e-4 at case1 "
                                         "(inside the 1e-3 contract, outside it for case3/case4)"}
    Td, Pd = 64, 16384
    cd, ld = synthetic_inputs(cin, L, Td, Pd)
    cd, ld = cd.to(dev)[None], ld.to(dev)
    mask = torch.zeros(Pd, 1, device=dev)
    mask[torch.randperm(Pd, device=dev)[:1000]] = 1.0
    y_meas = torch.randn(Td, Pd, cout, device=dev) * 0.05

    def dps_step():
        l = ld[:, None].detach().requires_grad_(True)
        y = model(cd, l)
        loss = torch.linalg.norm((y_meas - y) * mask)  # condition_methods.py:30-31
        torch.autograd.grad(loss, l)                    # condition_methods.py:32

    ms = timed(dps_step, 5)
    out["dps_fwd_bwd"] = {"value": Td * Pd / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                          "workload": f"{CASE} shapes, {Td} latents x {Pd} points, 1000 random sensors, "
                                      "forward(+cos stash) + loss + backward to dL/dlatent (BASELINE config 4)",
                          "precision": args.precision}

    # ---- case4 (3-D recipe: 15 hidden layers of width 384), the shapes of BASELINE configs 3 and 4, at a bounded size
    del y_meas, mask, cd, ld
    dims4 = CASE4_DIMS
    m4 = seeded_model(cb, dims4, args.precision).eval().to(dev)
    T4, P4 = 32, 131072
    c4, l4 = synthetic_inputs(dims4[0], dims4[1], T4, P4)
    c4, l4 = c4.to(dev)[None], l4.to(dev)
    with torch.no_grad():
        ms = timed(lambda: m4(c4, l4[:, None]), 3)
    flops4 = 2 * (dims4[0] * dims4[4] + dims4[3] * dims4[4] ** 2 + dims4[4] * dims4[2])
    out["case4_decode"] = {"value": T4 * P4 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                           "workload": f"case4 shapes {dict(zip(('cin', 'L', 'cout', 'nl', 'H'), dims4))}, {T4} frames x {P4} "
                                       "points, forward only (BASELINE config 3 shapes on one GPU, frames reduced)",
                           "precision": args.precision,
                           "algorithmic_tflops": T4 * P4 * flops4 / (ms * 1e-3) / 1e12}
    Td4, Pd4 = 64, 16384
    cd4 = c4[:, :Pd4].contiguous()
    ld4 = torch.cat([l4, l4], dim=0)[:Td4].contiguous()
    mask4 = torch.zeros(Pd4, 1, device=dev)
    mask4[torch.randperm(Pd4, device=dev)[:1000]] = 1.0
    y_meas4 = torch.randn(Td4, Pd4, dims4[2], device=dev) * 0.05

    def dps_step4():
        l = ld4[:, None].detach().requires_grad_(True)
        y = m4(cd4, l)
        loss = torch.linalg.norm((y_meas4 - y) * mask4)
        torch.autograd.grad(loss, l)

    ms = timed(dps_step4, 3)
    out["dps_fwd_bwd_case4"] = {"value": Td4 * Pd4 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                "workload": f"case4 shapes, {Td4} latents x {Pd4} points, 1000 random sensors, forward(+cos "
                                            "stash) + loss + backward to dL/dlatent (BASELINE config 4)",
                                "precision": args.precision}
    return out


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
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line (no NCCL version banner)
        dist.init_process_group("nccl", device_id=dev)

    cin, L, cout, nl, H = DIMS
    T, P = args.frames, args.points
    model = seeded_model(cb, DIMS, args.precision).eval().to(dev)
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
            gather_mode = "fused into the decode kernel: epilogue stores to all ranks' symmetric-memory buffers over NVLink"
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

    extra = {}
    if world == 1 and not args.no_extra:
        extra = measure_extra(model, coords, lat, dev, args)

    if rank == 0:
        peaks, peak_src = load_peaks()
        kavg_ms = sum(kern_ms) / max(1, len(kern_ms))
        fl = flops_per_pf(*DIMS) * T * P
        ach_tf = fl / (kavg_ms * 1e-3) / 1e12
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        sin_rate = sins_per_pf(*DIMS) * T * P / (kavg_ms * 1e-3)
        mufu_peak = MUFU_PEAK_SIN_PER_S
        launch = _native.query_launch(model._cdims(), model._precision_code(), T, P)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"CNF decode {CASE} 2D shapes, {T} frames x {P} points per GPU, forward only "
                                   "(BASELINE.json configs[1])",
                       "dims": dict(zip(("cin", "L", "cout", "nl", "H"), DIMS)), "frames_per_gpu": T, "points": P,
                       "precision": {"bf16x3": "tcgen05 bf16 hi/lo split, 3 MMAs per product, fp32 accumulate",
                                     "fp16": "tcgen05 single fp16 MMA per product, fp32 accumulate",
                                     "fp32": "CUDA-core fp32 FMA"}[args.precision],
                       "parallelism": f"frames sharded over {world} GPU(s)" + (f"; all-gather of the field inside the step, {gather_mode}" if world > 1 else ""),
                       "l2": "256 MiB memset between timed steps; each step also writes %.0f MB of output" % (T * P * cout * 4 / 1e6)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(coords_h.numel() * 4 + lat_h.numel() * 4),
                    "d2h_bytes_per_step": int(out_h.numel() * 4), "steps": e2e_steps,
                    "api": "confild_b200.decoder(coords, latents, model, x_normalizer, y_normalizer, 16, device) "
                           "with pinned host buffers (mirror of cnf/inference_function.py:51-76)"},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                         "kernel": "tc2_forward_kernel" if args.precision != "fp32" else "simt_forward_kernel",
                         "kernel_ms": kavg_ms, "peak_source": f"{peak_src} bf16_tflops_sustained",
                         "algorithmic_flops_per_launch": fl,
                         "mufu": {"achieved_gsin_s": sin_rate / 1e9, "measured_peak_gsin_s": mufu_peak / 1e9,
                                  "frac": sin_rate / mufu_peak}},
            "launch": dict(zip(("sms", "ctas", "threads", "smem_bytes", "ctas_per_sm", "tmem_cols", "tile_points"), launch)),
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if extra:
            line["extra"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CONFILD_PRECISION", "bf16x3"), choices=["bf16x3", "fp16", "fp32"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--points", type=int, default=POINTS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the fast-mode and DPS side measurements")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="N>1: how the decoded field is all-gathered")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
