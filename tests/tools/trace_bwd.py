"""Debug: event trace of CTA 0 of tc2_backward_kernel (needs the -DCNF_TRACE build)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["CONFILD_CNF_LIB"] = os.path.join(ROOT, "confild_b200", "libconfild_cnf_trace.so")
import torch
import confild_b200 as cb
from confild_b200 import _native
from oracle import cnf_oracle as O
dims = O.CASE_SHAPES["case1"]; sd = O.init_params(*dims, seed=0)
T, P = 64, 16384
coords, lat = O.synthetic_inputs(dims[0], dims[1], T, P)
m = cb.SIRENAutodecoder_film(2, 128, 3, 10, 128, precision="bf16x3"); m.load_state_dict(sd); m = m.eval().cuda()
c = coords.cuda()[None]; gout = torch.randn(T, P, 3, device="cuda")
def step():
    l = lat.cuda()[:, None].requires_grad_(True)
    y = m(c, l)
    return y, l
y, l = step(); torch.autograd.grad(y, l, grad_outputs=gout); torch.cuda.synchronize()
y, l = step(); torch.cuda.synchronize()
buf = torch.zeros(24 * 8192, dtype=torch.int64, device="cuda")
lib = _native.load()
assert lib.cnf_debug_set_trace(ctypes.c_void_p(buf.data_ptr())) == 0
torch.autograd.grad(y, l, grad_outputs=gout); torch.cuda.synchronize()
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
for l in (7, 6, 5):
    print(f"--- tile-pair #{TILE}, layer {l}")
    for g in (0, 1):
        r = nth(ev[2 + g], 2000 + l, TILE); i = nth(ev[2 + g], 3000 + l, TILE)
        print(f"  slot {g}: operands ready {r}, issued {i} (+{i - r})")
        for w in (8 * g, 8 * g + 4):
            p0 = nth(ev[4 + w], 200 + l, TILE); d = nth(ev[4 + w], 300 + l, TILE); e = nth(ev[4 + w], 400 + l, TILE)
            print(f"    warp {w:2d}: prefetch issued {p0}, d_full seen {d} (+{d - i} after issue), epilogue done {e} (E={e - d})")
