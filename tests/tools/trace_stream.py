"""Debug: per-episode event trace of the issue stream of tc_forward_kernel (H = 256 / 384), CTA 0, -DCNF_TRACE build:
    CNF_TC_CLUSTER=2 python tests/tools/trace_stream.py case4 f16f8 [first_layer] [layers]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["CONFILD_CNF_LIB"] = os.path.join(ROOT, "confild_b200", "libconfild_cnf_trace.so")
import torch
import confild_b200 as cb
from confild_b200 import _native
from oracle import cnf_oracle as O
case = sys.argv[1] if len(sys.argv) > 1 else "case4"
prec = sys.argv[2] if len(sys.argv) > 2 else "f16f8"
L0 = int(sys.argv[3]) if len(sys.argv) > 3 else 40
NL = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dims = O.CASE_SHAPES[case]; sd = O.init_params(*dims, seed=0)
coords, lat = O.synthetic_inputs(dims[0], dims[1], 16, 16384)
m = cb.SIRENAutodecoder_film(*dims[:2], dims[2], dims[3], dims[4], precision=prec); m.load_state_dict(sd); m = m.eval().cuda()
c, l = coords.cuda()[None], lat.cuda()[:, None]
with torch.no_grad():
    m(c, l); torch.cuda.synchronize()
    buf = torch.zeros(24 * 8192, dtype=torch.int64, device="cuda")
    lib = _native.load()
    assert lib.cnf_debug_set_trace(ctypes.c_void_p(buf.data_ptr())) == 0
    m(c, l); torch.cuda.synchronize()
b = buf.cpu().view(24, 4096, 2)
t0 = int(b[b[:, :, 1] > 0][:, 1].min())
NB = dims[4] // 128
kE = NB * (dims[4] // 64)
ep = {}
for w in range(3):
    for code, t in b[20 + w]:
        code, t = int(code), int(t)
        if t <= 0: continue
        kind, rest = divmod(code, 100000)
        layer, idx = divmod(rest, 100)
        ep.setdefault((layer, idx), {})[kind] = t - t0
        ep[(layer, idx)]["w"] = w
print(f"{case} {prec} cluster={os.environ.get('CNF_TC_CLUSTER', '1')}: episodes of global layers {L0}..{L0 + NL - 1}  "
      "(times in clk; wait = start, +A = operand/accumulator ready, +B = stages landed, +tok = token, +iss = issued)")
# epilogue events of activation warp 0 (role 4) and warp 15 (role 19): 300+10l+n block n seen complete, 600+10l+n epilogue done
epi = []
for role in (4, 19):
    for code, t in b[role]:
        code, t = int(code), int(t)
        if t > 0 and (300 <= code < 500 or 600 <= code < 800):
            k = "seen" if code < 500 else "done"
            c = code - (300 if code < 500 else 600)
            epi.append((t - t0, f"    epilogue warp {role - 4:2d}: block {c % 10} of tile-layer {c // 10} {k}"))
epi.sort()
win_lo = ep[(L0, 0)][1]; win_hi = ep[(L0 + NL - 1, kE - 1)][5]
lines = []
prev_iss = None
for layer in range(L0, L0 + NL):
    for idx in range(kE):
        e = ep.get((layer, idx))
        if not e or 5 not in e: continue
        kp = idx // (2 * NB); r = idx - kp * 2 * NB; n = r >> 1; ks = 2 * kp + (r & 1)
        gap = "" if prev_iss is None else f"  since prev issued {e[5] - prev_iss:6d}"
        lines.append((e[5], f"L{layer} ep{idx:2d} (block {n}, slab {ks}) w{e['w']}: wait {e[1]:8d}  +A {e[2] - e[1]:6d}  +B {e[3] - e[2]:6d}  "
                      f"+tok {e[4] - e[3]:6d}  +iss {e[5] - e[4]:6d} = issued {e[5]:8d}{gap}"))
        prev_iss = e[5]
lines += [(t, txt + f" at {t}") for t, txt in epi if win_lo <= t <= win_hi]
for _, txt in sorted(lines):
    print(txt)
firsts = [ep[(la, 0)][5] for la in range(L0 - 8, L0 + 8) if (la, 0) in ep and 5 in ep[(la, 0)]]
print("layer period (ep0 issued):", [b_ - a_ for a_, b_ in zip(firsts, firsts[1:])])
