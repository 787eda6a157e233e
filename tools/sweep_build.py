"""Builds tuning variants of libb2exec.so into build/variants/ (they travel to
the GPU box with the snapshot).  Usage: python tools/sweep_build.py"""
import itertools
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from covalent_ssh_plugin_b200.build import build_variant  # noqa: E402

out = REPO / "build" / "variants"
for threads, ctas, vec in [(256, 4, 4), (128, 8, 4), (128, 6, 4), (128, 8, 8), (64, 12, 4), (64, 16, 4), (64, 12, 8)]:
    p = build_variant(out / f"t{threads}_w{ctas}_v{vec}.so",
                      {"B2_WORKER_THREADS": threads, "B2_WORKER_MIN_CTAS": ctas, "B2_WORKER_VEC": vec})
    print(p)
