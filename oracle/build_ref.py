"""Builds `oracle/_ref/` from the reference sources WHERE THEY LIE — TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            (run by `__graft_entry__.build()` when /root/reference exists)

The reference is Python: "compiling" it means byte-compiling `src/tinyedm/networks.py` and `src/tinyedm/solvers.py`
(the two files of the hot path that import with torch + numpy alone — `import tinyedm` itself needs lightning /
torchmetrics / hydra / diffusers, absent here) into `oracle/_ref/*.code` (marshalled code objects). No reference SOURCE enters the repository:
`oracle/_ref/` is git-ignored (it is not gpurun-ignored, so the bytecode travels to the GPU box like the built `.so`).
Consumers: `oracle/ref_loader.py` -> `bench.py --impl reference` / `cpu_baseline` (kind "reference") and the `-m gpu`
noise-floor tests, which run the reference's own modules under bf16 autocast on the B200. The product
package never imports any of it.
"""
from __future__ import annotations

import marshal
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TINYEDM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ("networks", "solvers")


def build() -> bool:
    src_dir = os.path.join(REF, "src", "tinyedm")
    if not all(os.path.exists(os.path.join(src_dir, f + ".py")) for f in FILES):
        return False
    os.makedirs(OUT, exist_ok=True)
    for f in FILES:
        # (a marshalled code object under a neutral extension: snapshot tools commonly skip *.pyc)
        with open(os.path.join(src_dir, f + ".py")) as fh:
            code = compile(fh.read(), f"reference/src/tinyedm/{f}.py", "exec", dont_inherit=True)
        with open(os.path.join(OUT, f + ".code"), "wb") as fh:
            fh.write(marshal.dumps(code))
    with open(os.path.join(OUT, "README"), "w") as fh:
        fh.write(f"bytecode of {REF}/src/tinyedm/{{networks,solvers}}.py, python {sys.version.split()[0]}; built by oracle/build_ref.py\n")
    return True


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref built" if ok else f"reference sources not found under {REF}: nothing built")
