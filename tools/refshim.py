"""Import the UNMODIFIED reference (Deep-MI/LaPy) from a scratch copy - dev container only.

Used by tools/make_golden.py and tools/* probes.  Never imported by the product, the tests,
bench.py or smoke(): /root/reference does not exist on the GPU box.  The copy protects the
read-only reference tree (its own test-suite writes into data/) and the
``importlib.metadata.version`` patch works around lapy/_version.py:5.
"""

import importlib.metadata as _md
import os
import shutil
import sys

REF = "/root/reference"
SCRATCH = "/tmp/lapy_ref_copy"


def load():
    if not os.path.isdir(os.path.join(SCRATCH, "lapy")):
        os.makedirs(SCRATCH, exist_ok=True)
        shutil.copytree(os.path.join(REF, "lapy"), os.path.join(SCRATCH, "lapy"))
        shutil.copytree(os.path.join(REF, "data"), os.path.join(SCRATCH, "data"))
    _orig = _md.version
    _md.version = lambda n: "1.6.0.dev0" if n == "lapy" else _orig(n)
    if SCRATCH not in sys.path:
        sys.path.insert(0, SCRATCH)
    import lapy  # noqa: E402

    return lapy, os.path.join(SCRATCH, "data")
