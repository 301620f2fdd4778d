"""Debug: event trace of CTA 0 of the generic tc_forward_kernel (needs the -DCNF_TRACE build):
    python tests/tools/trace_generic.py case4 bf16x3"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["CONFILD_CNF_LIB"] = os.path.join(ROOT, "confild_b200", "libconfild_cnf_trace.so")
import torch
import confild_b200 as cb
from confild_b200 import _native
from oracle import cnf_oracle as O
case = sys.argv[1] if len(sys.argv) > 1 else "case4"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
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
ev = {r: [(int(c), int(t) - t0) for c, t in b[r] if t > 0] for r in range(24)}
def nth(evs, code, n):
    k = 0
    for c, t in evs:
        if c == code:
            if k == n: return t
            k += 1
    return None
TILE = 3
NB = dims[4] // 128
prev = None
for l in (3, 4, 5, 6):
    parts = []
    for n in range(NB):
        d = nth(ev[4], 300 + 10 * l + n, TILE); e = [nth(ev[4 + w], 600 + 10 * l + n, TILE) for w in (0, 5, 10, 15)]
        parts.append(f"block {n}: complete seen {d}, epilogue done +{[x - d for x in e]}")
    iss = [nth(ev[2 + n], 3000 + l, TILE) for n in range(NB)]
    d0 = nth(ev[4], 300 + 10 * l, TILE)
    print(f"layer {l}: " + " | ".join(parts) + f" | issuers done {iss}" + (f" | period {d0 - prev}" if prev else ""))
    prev = d0
