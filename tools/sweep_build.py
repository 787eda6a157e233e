"""Builds tuning variants of libb2exec.so into build/variants/ (they travel to
the GPU box with the snapshot).  Usage: python tools/sweep_build.py"""
import itertools
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from covalent_ssh_plugin_b200.build import build_variant  # noqa: E402

out = REPO / "build" / "variants"
for threads, ctas, stages, tile in [(256, 4, 3, 16384), (256, 3, 4, 16384), (256, 2, 3, 32768), (128, 4, 3, 16384),
                                    (512, 2, 6, 16384), (256, 3, 2, 32768)]:
    p = build_variant(out / f"t{threads}_w{ctas}_s{stages}_tile{tile}.so",
                      {"B2_WORKER_THREADS": threads, "B2_WORKER_MIN_CTAS": ctas, "B2_TMA_STAGES": stages,
                       "B2_TMA_TILE": tile})
    print(p)
