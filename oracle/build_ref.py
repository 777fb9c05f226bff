"""Recipe for `oracle/_ref/`: byte-compiles the reference's own hot-path modules where they lie under /root/reference
into `oracle/_ref/strhub/**.pyc` (binaries only — no reference source enters the repo; `oracle/_ref/` is git-ignored but
travels to the GPU box with the snapshot, like the built `.so`).  With it `bench.py --impl reference` and the
`cpu_baseline` leg time the UNMODIFIED reference `strhub.models.parseq.model.PARSeq` (under oracle/timm_shim.py for the
three timm names) on the GPU box's host cores instead of the oracle port.

    python -m oracle.build_ref

TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under parseq_b200/ imports it.
"""
from __future__ import annotations

import os
import py_compile
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("PARSEQ_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")

# strhub/models/parseq/model.py:23-28 import graph (timm comes from oracle/timm_shim.py)
FILES = [
    "strhub/__init__.py",
    "strhub/data/__init__.py",
    "strhub/data/utils.py",                 # Tokenizer
    "strhub/models/__init__.py",
    "strhub/models/utils.py",               # init_weights
    "strhub/models/parseq/__init__.py",
    "strhub/models/parseq/model.py",        # PARSeq.forward / encode / decode
    "strhub/models/parseq/modules.py",      # Encoder, Decoder, DecoderLayer, TokenEmbedding
]


def build(verbose: bool = False) -> bool:
    """Returns True when oracle/_ref is usable afterwards (freshly built, or already present)."""
    if not os.path.isfile(os.path.join(SRC, FILES[-1])):
        return os.path.isfile(os.path.join(OUT, "strhub/models/parseq/model.pyc"))
    for rel in FILES:
        dst = os.path.join(OUT, rel[:-3] + ".pyc")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path recorded in tracebacks (the reference's own), sourceless import does not need the source
        py_compile.compile(os.path.join(SRC, rel), cfile=dst, dfile=os.path.join("/root/reference", rel), doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        if verbose:
            print("compiled", rel, "->", os.path.relpath(dst, ROOT))
    with open(os.path.join(OUT, "README"), "w") as f:
        f.write("byte-compiled from /root/reference by oracle/build_ref.py (python %d.%d); not source, not tracked\n"
                % sys.version_info[:2])
    return True


if __name__ == "__main__":
    print("oracle/_ref ready:", build(verbose=True))
