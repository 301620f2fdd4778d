"""Debug: event trace of CTA 0 of tc3_forward_kernel (needs the -DCNF_TRACE build; CNF_TC3=1 is set here)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["CONFILD_CNF_LIB"] = os.path.join(ROOT, "confild_b200", "libconfild_cnf_trace.so")
os.environ["CNF_TC3"] = "1"
import torch
import confild_b200 as cb
from confild_b200 import _native
from oracle import cnf_oracle as O
dims = O.CASE_SHAPES["case1"]; sd = O.init_params(*dims, seed=0)
coords, lat = O.synthetic_inputs(dims[0], dims[1], 8, 65536)
m = cb.SIRENAutodecoder_film(2, 128, 3, 10, 128, precision="f16f8"); m.load_state_dict(sd); m = m.eval().cuda()
c, l = coords.cuda()[None], lat.cuda()[:, None]
with torch.no_grad():
    m(c, l); torch.cuda.synchronize()
    buf = torch.zeros(24 * 8192, dtype=torch.int64, device="cuda")
    lib = _native.load()
    assert lib.cnf_debug_set_trace(ctypes.c_void_p(buf.data_ptr())) == 0
    m(c, l); torch.cuda.synchronize()
b = buf.cpu().view(24, 4096, 2)
t0 = int(b[b[:, :, 1] > 0][:, 1].min())
rows = []
names = {2: "W0", 3: "W1", 4: "T0", 12: "T1"}
for role, nm in names.items():
    for code, t in b[role]:
        if t > 0:
            rows.append((int(t) - t0, nm, int(code)))
rows.sort()
# print a window in the steady state: the third trip of CTA 0
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
hi = int(sys.argv[2]) if len(sys.argv) > 2 else lo + 40000
kind = {0: "wB", 1: "wait", 2: "ready", 3: "issued", 4: "done"}
for t, nm, code in rows:
    if lo <= t < hi:
        k, r = divmod(code, 1000)
        if k == 0: desc = f"wait-weights l={r - 500}"
        else: desc = f"{kind[k]:6s} l={r // 10} g={r % 10}"
        col = {"W0": 0, "W1": 1, "T0": 2, "T1": 3}[nm]
        print(f"{t:8d} " + " " * (26 * col) + f"{nm} {desc}")
print("total span", rows[-1][0])
