"""Generates tests/golden/golden_v1.json by running every case of cases.py through
the UNMODIFIED reference plugin (oracle/ref_harness.py: the reference's ssh.py +
exec.py with only the SSH byte transport replaced by local file copies).

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py
"""

from __future__ import annotations

import hashlib
import json
import sys
import tempfile
from pathlib import Path

import cloudpickle
import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))

from oracle import ref_harness  # noqa: E402
from tests.golden import cases  # noqa: E402

cloudpickle.register_pickle_by_value(cases)


def describe(value):
    """JSON description of a result object, exact to the bit."""
    if isinstance(value, np.ndarray):
        raw = value.tobytes(order="A")
        d = {"type": "ndarray", "dtype": str(value.dtype), "shape": list(value.shape),
             "c_contiguous": bool(value.flags.c_contiguous), "f_contiguous": bool(value.flags.f_contiguous),
             "sha256": hashlib.sha256(raw).hexdigest(), "nbytes": len(raw)}
        if len(raw) <= 64:
            d["hex"] = raw.hex()
        return d
    if isinstance(value, np.generic):
        return {"type": "numpy_scalar", "dtype": str(value.dtype), "hex": value.tobytes().hex()}
    if type(value) is int:
        return {"type": "int", "value": str(value)}
    if isinstance(value, str):
        return {"type": "str", "value": value}
    raise TypeError(f"no description for {type(value)}")


def main():
    if not ref_harness.available():
        raise SystemExit("the reference is not present; golden vectors can only be made in the build container")
    out = {"generator": "tests/golden/make_golden.py", "reference": "covalent-ssh-plugin 0.24.1 (unmodified ssh.py + exec.py, local transport)",
           "numpy": np.__version__, "cloudpickle": cloudpickle.__version__, "cases": {}}
    with tempfile.TemporaryDirectory() as tmp:
        ex = ref_harness.make_executor(tmp, pythonpath=str(REPO))
        for i, (cid, ename, recipes) in enumerate(cases.CASES):
            fn = cases.ELECTRONS[ename]
            args = cases.build_args(recipes)
            try:
                res = ref_harness.reference_run(ex, fn, args, {}, node_id=i)
                out["cases"][cid] = {"electron": ename, "result": describe(res)}
            except Exception as err:
                out["cases"][cid] = {"electron": ename,
                                     "raises": {"type": type(err).__name__, "args": [repr(a) for a in err.args]}}
            print(cid, "->", json.dumps(out["cases"][cid])[:140])
    path = Path(__file__).with_name("golden_v1.json")
    path.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", path, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
