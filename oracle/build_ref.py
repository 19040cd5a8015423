"""Recipe: make the UNMODIFIED reference's hot-path files importable on the GPU box.  TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py [--ref /root/reference]

The reference is a Python tree; /root/reference exists only in the build container.  This script copies — byte for
byte, at build time, never into git history — the few files that make up the path

    models/cdan.py  models/cbam.py  utils/post_processing.py  utils/postprocessing_factory.py

into the git-ignored directory ``oracle/_ref/`` (which is NOT gpurun-ignored, so it travels with the snapshot like the
built ``.so``), together with a MANIFEST of sha256 sums.  ``load_ref()`` imports them from there under private module
names handling (the reference uses absolute imports ``from models.cbam import CBAM`` that collide with the drop-in
package's own ``models``), so that

  * ``bench.py --impl reference`` and ``cpu_baseline`` time the reference's own ``CDAN.forward``
    (/root/reference/models/cdan.py:171-176) on the box's host cores (kind = "reference"),
  * ``-m gpu`` parity tests compare the CUDA path with the reference itself, not only with the restatement.

Nothing in the product imports this module.  If ``oracle/_ref`` is absent (a checkout that never ran build()), callers
fall back to the oracle port and say so (kind = "port").
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ["models/cdan.py", "models/cbam.py", "utils/post_processing.py", "utils/postprocessing_factory.py"]


def build_ref(ref_root: str = "/root/reference") -> bool:
    """Copy the reference's path files into oracle/_ref (returns False when the reference tree is not present)."""
    if not all(os.path.isfile(os.path.join(ref_root, f)) for f in FILES):
        return False
    lines = []
    for f in FILES:
        dst = os.path.join(REF_DIR, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(ref_root, f), dst)
        with open(dst, "rb") as fh:
            lines.append(f"{hashlib.sha256(fh.read()).hexdigest()}  {f}")
    with open(os.path.join(REF_DIR, "MANIFEST"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return True


def ref_available() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, f)) for f in FILES)


def _clashing(name: str) -> bool:
    return name in ("models", "utils") or name.startswith("models.") or name.startswith("utils.")


def load_ref(ref_root: str = REF_DIR):
    """Import (cdan_module, post_processing_module, postprocessing_factory_module) of the reference from `ref_root`
    without leaving its ``models`` / ``utils`` packages in sys.modules (the drop-in package uses the same names)."""
    saved = {k: v for k, v in sys.modules.items() if _clashing(k)}
    for k in saved:
        del sys.modules[k]
    saved_path = list(sys.path)
    sys.path[:] = [ref_root] + [p for p in sys.path if "multi-degradation-image-enhancement_b200" not in p]
    importlib.invalidate_caches()
    try:
        cdan_mod = importlib.import_module("models.cdan")
        pp_mod = importlib.import_module("utils.post_processing")
        ppf_mod = importlib.import_module("utils.postprocessing_factory")
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if _clashing(k)]:
            del sys.modules[k]
        sys.modules.update(saved)
    return cdan_mod, pp_mod, ppf_mod


def reference_forward_fn(sd):
    """(callable x -> y, kind): the reference's own CDAN in eval mode when oracle/_ref exists, else the oracle port."""
    import torch
    if ref_available():
        cdan_mod, _, _ = load_ref()
        net = cdan_mod.CDAN()
        net.load_state_dict(sd, strict=True)
        net.eval()

        def run(x):
            with torch.no_grad():
                return net(x)
        return run, "reference"
    from oracle.cdan_oracle import cdan_forward

    def run_port(x):
        with torch.no_grad():
            return cdan_forward(sd, x)
    return run_port, "port"


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    ok = build_ref(a.ref)
    print("oracle/_ref written" if ok else f"reference tree not found at {a.ref}; oracle/_ref unchanged")
