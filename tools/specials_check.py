#!/usr/bin/env python
"""IEEE-754 corner-value golden cases on the GPU, with a diagnosis instead of an assert: for every case whose
bytes differ from the reference's, the first few differing elements (operand bits, device bits, numpy bits)."""
import asyncio
import hashlib
import json
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import numpy as np  # noqa: E402

from tests.golden import cases  # noqa: E402


def main():
    from covalent_ssh_plugin_b200 import SSHExecutor

    golden = json.loads((Path(__file__).resolve().parents[1] / "tests/golden/golden_v1.json").read_text())["cases"]
    tmp = tempfile.mkdtemp(prefix="b2-spec-")
    key = Path(tmp) / "id"
    key.touch()
    ex = SSHExecutor(username="u", hostname="cuda:0", ssh_key_file=str(key), cache_dir=tmp)
    todo = [c for c in cases.CASES if "specials" in c[0]]
    bad = 0
    for i, (cid, ename, recipes) in enumerate(todo):
        args = cases.build_args(recipes)
        got = asyncio.run(ex.run(cases.ELECTRONS[ename], args, {}, {"dispatch_id": "spec", "node_id": i}))
        with np.errstate(all="ignore"):
            want = cases.ELECTRONS[ename](*args)
        sha = hashlib.sha256(got.tobytes(order="A")).hexdigest()
        ok = sha == golden[cid]["result"]["sha256"]
        gb, wb = got.view(np.uint32).ravel(), want.view(np.uint32).ravel()
        diff = np.flatnonzero(gb != wb)
        print(f"{cid}: {'bit-identical to the reference' if ok else 'DIFFERS'}; vs in-process numpy: {len(diff)} of {gb.size} elements differ")
        for j in diff[:8]:
            ops = " ".join(f"{a.view(np.uint32).ravel()[j]:08x}" for a in args)
            print(f"    [{j}] operands {ops} -> device {gb[j]:08x}, numpy {wb[j]:08x}")
        bad += not ok
    print("specials:", "all bit-identical" if not bad else f"{bad} case(s) differ")


if __name__ == "__main__":
    main()
