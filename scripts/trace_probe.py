"""Debug: event trace of CTA 0 of tc2_forward_kernel (needs the -DCNF_TRACE build)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
    buf = torch.zeros(6 * 8192, dtype=torch.int64, device="cuda")
    lib = _native.load()
    assert lib.cnf_debug_set_trace(ctypes.c_void_p(buf.data_ptr())) == 0
    m(c, l); torch.cuda.synchronize()
b = buf.cpu().view(6, 4096, 2)
t0 = int(b[b[:, :, 1] > 0][:, 1].min())
for role, name in enumerate(["WG0", "WG1", "MMA", "PROD", "WG0hf1", "WG1hf1"]):
    ev = [(int(c), int(t) - t0) for c, t in b[role] if t > 0]
    # skip first tile (cold), print second tile-pair
    print(name, "events", len(ev))
    start = None
    shown = 0
    for i, (c, t) in enumerate(ev):
        if role in (0, 1, 4, 5) and c == 100:
            shown += 1
        if role == 2 and c % 100 == 1 and 1000 <= c < 1100:
            shown += 1
        if role == 3 and c == 6010:
            shown += 1
        if shown == 3:
            print(f"  {c:5d} t={t:8d}" + (f"  (+{t - ev[i-1][1]})" if i else ""))
        if shown > 3:
            break
