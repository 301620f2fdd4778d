"""Recipe: byte-compile the reference's own CNF decoder into ``oracle/_ref/`` (git-ignored build output).

    python oracle/build_ref.py            # build container only: needs /root/reference

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is pure Python, so its "build" is ``py_compile``: the three
modules on the decode path (ConditionalNeuralField/cnf/{nf_networks,components,initialization}.py) are compiled
from where they lie under /root/reference into bytecode files ``oracle/_ref/ConditionalNeuralField/cnf/<module>.bin``
(the ``.pyc`` format under a neutral suffix: snapshot tools commonly drop ``*.pyc``), loaded here by un-marshalling
the code objects and executing them as modules named ``ConditionalNeuralField.cnf.<module>``.  No reference source is copied into the repository and ``oracle/_ref/``
stays out of git history; like a compiled ``.so`` it travels with the snapshot to the GPU box, where
``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference's OWN ``SIRENAutodecoder_film.forward``
on the host cores (``cpu_baseline.kind = "reference"``).  When ``oracle/_ref`` is absent they fall back to the
restated port (``oracle/cnf_oracle.py``, ``kind = "port"``).  The product path never imports either.
"""
from __future__ import annotations

import importlib.util
import marshal
import os
import py_compile
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
OUT_ROOT = os.path.join(HERE, "_ref")
PKG = os.path.join("ConditionalNeuralField", "cnf")
MODULES = ["nf_networks", "components", "initialization"]


MODULES_IN_IMPORT_ORDER = ["initialization", "components", "nf_networks"]


def _bin(m: str) -> str:
    return os.path.join(OUT_ROOT, PKG, m + ".bin")


def available() -> bool:
    return all(os.path.exists(_bin(m)) for m in MODULES)


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds the compiled reference (built now or earlier)."""
    src_dir = os.path.join(REF_ROOT, PKG)
    if not os.path.isdir(src_dir):
        return available()
    os.makedirs(os.path.join(OUT_ROOT, PKG), exist_ok=True)
    for m in MODULES:
        src, dst = os.path.join(src_dir, m + ".py"), _bin(m)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=f"<reference>/{PKG}/{m}.py", doraise=True)
    return available()


def load_reference_class():
    """The reference's own ``SIRENAutodecoder_film`` from the compiled modules, or None when oracle/_ref is absent."""
    if not available():
        return None
    if "ConditionalNeuralField.cnf.nf_networks" not in sys.modules:
        for pkg in ("ConditionalNeuralField", "ConditionalNeuralField.cnf"):
            if pkg not in sys.modules:
                mod = types.ModuleType(pkg)
                mod.__path__ = []  # a package with no importable source: its modules are registered below
                sys.modules[pkg] = mod
        sys.modules["ConditionalNeuralField"].cnf = sys.modules["ConditionalNeuralField.cnf"]
        for m in MODULES_IN_IMPORT_ORDER:  # dependencies first: the modules import one another by absolute name
            with open(_bin(m), "rb") as f:
                blob = f.read()
            if blob[:4] != importlib.util.MAGIC_NUMBER:
                raise RuntimeError(f"{_bin(m)} was compiled by another Python version; rebuild with oracle/build_ref.py")
            code = marshal.loads(blob[16:])  # .pyc layout: 16-byte header, then the marshalled module code object
            name = f"ConditionalNeuralField.cnf.{m}"
            mod = types.ModuleType(name)
            mod.__file__ = _bin(m)
            mod.__package__ = "ConditionalNeuralField.cnf"
            sys.modules[name] = mod
            setattr(sys.modules["ConditionalNeuralField.cnf"], m, mod)
            exec(code, mod.__dict__)
    return sys.modules["ConditionalNeuralField.cnf.nf_networks"].SIRENAutodecoder_film


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference and no earlier build)")
