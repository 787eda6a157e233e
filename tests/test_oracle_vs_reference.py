"""Validates the oracle port against the REAL reference where it is present
(/root/reference exists only in the build container; skipped elsewhere): the
same random electrons through the unmodified ssh.py/exec.py (oracle/ref_harness.py)
and through oracle/reference_port.py must agree bit for bit, including the
artefact names the reference leaves behind with do_cleanup=False."""

import asyncio
import os
from pathlib import Path

import numpy as np
import pytest

from oracle import ref_harness as R
from oracle import reference_port as P

REPO = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.skipif(not R.available(), reason="reference sources not present on this box")


def test_port_equals_reference_on_random_electrons(tmp_path):
    from covalent_ssh_plugin_b200 import electrons as E

    rng = np.random.default_rng(2024)
    x = rng.standard_normal(4097, dtype=np.float32)
    y = rng.standard_normal(4097, dtype=np.float32)
    calls = [(E.add_int, [int(rng.integers(-10 ** 12, 10 ** 12)), 17]), (E.mul_int, [1 << 40, 1 << 40]),
             (E.scale_add, [x]), (E.vec_mul, [x, y]), (E.vec_sum, [np.abs(x)]),
             (E.stack_sum, [[x, y, x]])]
    ref_ex = R.make_executor(str(tmp_path / "ref"), pythonpath=str(REPO))
    s = P.PortSettings(str(tmp_path / "port"), pythonpath=str(REPO))
    for i, (fn, args) in enumerate(calls):
        a = R.reference_run(ref_ex, fn, args, {}, node_id=i)
        b = asyncio.run(P.run_electron(s, fn, list(args), {}, {"dispatch_id": "golden", "node_id": i}))
        assert type(a) is type(b)
        if isinstance(a, np.ndarray):
            assert a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()
        elif isinstance(a, np.generic):
            assert a.tobytes() == b.tobytes()
        else:
            assert a == b


def test_artefact_names_match_reference(tmp_path):
    ref_ex = R.make_executor(str(tmp_path / "ref"), pythonpath=str(REPO), do_cleanup=False)
    s = P.PortSettings(str(tmp_path / "port"), pythonpath=str(REPO), do_cleanup=False)
    R.reference_run(ref_ex, len, [[1, 2, 3]], {}, dispatch_id="dx", node_id=9)
    asyncio.run(P.run_electron(s, len, [[1, 2, 3]], {}, {"dispatch_id": "dx", "node_id": 9}))
    for sub in ("cache", "remote"):
        assert sorted(os.listdir(tmp_path / "ref" / sub)) == sorted(os.listdir(tmp_path / "port" / sub))
    assert sorted(os.listdir(tmp_path / "port" / "cache")) == ["exec_dx_9.py", "function_dx_9.pkl", "result_dx_9.pkl"]


def test_exception_is_reraised_identically(tmp_path):
    def boom(k):
        raise KeyError(k, "extra")

    ref_ex = R.make_executor(str(tmp_path / "ref"), pythonpath=str(REPO))
    s = P.PortSettings(str(tmp_path / "port"), pythonpath=str(REPO))
    errs = []
    for run in (lambda: R.reference_run(ref_ex, boom, ["k"], {}),
                lambda: asyncio.run(P.run_electron(s, boom, ["k"], {}, {"dispatch_id": "e", "node_id": 0}))):
        with pytest.raises(KeyError) as ei:
            run()
        errs.append(ei.value.args)
    assert errs[0] == errs[1] == ("k", "extra")
