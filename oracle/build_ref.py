"""Recipe: byte-compile the reference's own CNF decoder into ``oracle/_ref/`` (git-ignored build output).

    python oracle/build_ref.py            # build container only: needs /root/reference

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is pure Python, so its "build" is ``py_compile``: the three
modules on the decode path (ConditionalNeuralField/cnf/{nf_networks,components,initialization}.py) are compiled
from where they lie under /root/reference into sourceless ``.pyc`` files under
``oracle/_ref/ConditionalNeuralField/cnf/``.  No reference source is copied into the repository and ``oracle/_ref/``
stays out of git history; like a compiled ``.so`` it travels with the snapshot to the GPU box, where
``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference's OWN ``SIRENAutodecoder_film.forward``
on the host cores (``cpu_baseline.kind = "reference"``).  When ``oracle/_ref`` is absent they fall back to the
restated port (``oracle/cnf_oracle.py``, ``kind = "port"``).  The product path never imports either.
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
OUT_ROOT = os.path.join(HERE, "_ref")
PKG = os.path.join("ConditionalNeuralField", "cnf")
MODULES = ["nf_networks", "components", "initialization"]


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT_ROOT, PKG, m + ".pyc")) for m in MODULES)


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds the compiled reference (built now or earlier)."""
    src_dir = os.path.join(REF_ROOT, PKG)
    if not os.path.isdir(src_dir):
        return available()
    os.makedirs(os.path.join(OUT_ROOT, PKG), exist_ok=True)
    for m in MODULES:
        src, dst = os.path.join(src_dir, m + ".py"), os.path.join(OUT_ROOT, PKG, m + ".pyc")
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=f"<reference>/{PKG}/{m}.py", doraise=True)
    return available()


def load_reference_class():
    """The reference's own ``SIRENAutodecoder_film`` from the compiled modules, or None when oracle/_ref is absent."""
    if not available():
        return None
    if OUT_ROOT not in sys.path:
        sys.path.insert(0, OUT_ROOT)  # ConditionalNeuralField / cnf resolve as namespace packages of .pyc modules
    from ConditionalNeuralField.cnf.nf_networks import SIRENAutodecoder_film  # type: ignore

    return SIRENAutodecoder_film


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference and no earlier build)")
