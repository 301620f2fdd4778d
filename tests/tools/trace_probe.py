"""Debug: event trace of CTA 0 of tc2_forward_kernel (needs the -DCNF_TRACE build)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["CONFILD_CNF_LIB"] = os.path.join(ROOT, "confild_b200", "libconfild_cnf_trace.so")
import torch
import confild_b200 as cb
from confild_b200 import _native
from oracle import cnf_oracle as O
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
dims = O.CASE_SHAPES["case1"]; sd = O.init_params(*dims, seed=0)
coords, lat = O.synthetic_inputs(dims[0], dims[1], 8, 65536)
m = cb.SIRENAutodecoder_film(2,128,3,10,128, precision=prec); m.load_state_dict(sd); m = m.eval().cuda()
c, l = coords.cuda()[None], lat.cuda()[:, None]
with torch.no_grad():
    m(c, l); torch.cuda.synchronize()
    buf = torch.zeros(20 * 8192, dtype=torch.int64, device="cuda")
    lib = _native.load()
    assert lib.cnf_debug_set_trace(ctypes.c_void_p(buf.data_ptr())) == 0
    m(c, l); torch.cuda.synchronize()
b = buf.cpu().view(20, 4096, 2)
t0 = int(b[b[:, :, 1] > 0][:, 1].min())
import collections
ev = {}
for role in range(20):
    ev[role] = [(int(c), int(t) - t0) for c, t in b[role] if t > 0]
# issuer warp of K slab h = role 2 + h (events 1000/2000/3000 + 500*slot + layer); epilogue warps = role 4 + warp
def nth(evs, code, n):
    k = 0
    for c, t in evs:
        if c == code:
            if k == n: return t
            k += 1
    return None
TILE = 3
for l in (3, 4, 5):
    print(f"--- tile-pair #{TILE}, layer {l}")
    for g in (0, 1):
        row = []
        for h in (0, 1):
            w = nth(ev[2 + h], 1000 + 500 * g + l, TILE); r = nth(ev[2 + h], 2000 + 500 * g + l, TILE)
            i = nth(ev[2 + h], 3000 + 500 * g + l, TILE)
            row.append(f"half{h} wait {w} ready {r} issued {i}")
            last = i
        print(f"  slot {g}: " + " | ".join(row))
        for w in range(8 * g, 8 * g + 8, 4):
            d = nth(ev[4 + w], 300 + l, TILE); e = nth(ev[4 + w], 400 + l, TILE); hh = nth(ev[4 + w], 350 + l, TILE)
            print(f"    warp {w:2d} (hf={(w // 4) % 2} wq={w % 4}): d_full seen {d} (+{d - last}), a_half arrived {hh} (+{hh - d}), epilogue done {e} (E={e - d})")

# steady-state layer period of the pair: time between successive layers' (slot 0, half 0) issue events
ts = [nth(ev[2], 3000 + l, TILE) for l in range(2, 10)]
print("slot0 half0 issued at", ts, "period", [b - a for a, b in zip(ts, ts[1:])])
for g in (0, 1):
    w = 8 * g
    ds = [nth(ev[4 + w], 300 + l, TILE) for l in range(2, 10)]
    es = [nth(ev[4 + w], 400 + l, TILE) for l in range(2, 10)]
    print(f"slot {g} warp {w}: d_full seen", ds, " E durations", [e - d for d, e in zip(ds, es)], " gap E_done->next d_full", [d2 - e for e, d2 in zip(es, ds[1:])])
