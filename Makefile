# Convenience targets; everything is plain Python underneath.
PY ?= python

build:            ## nvcc -gencode arch=compute_100a,code=sm_100a ... -> covalent_ssh_plugin_b200/libb2exec.so
	$(PY) -m covalent_ssh_plugin_b200.build

test-cpu: build   ## oracle vs golden vectors, host logic, ABI symbols, gloo ranks
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build   ## parity through the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

bench: build      ## electrons/s + roofline + cpu_baseline, one GPU
	$(PY) bench.py --gpus 1 --steps 20 --warmup 3

golden:           ## regenerate tests/golden/golden_v1.json from the unmodified reference (build container only)
	$(PY) tests/golden/make_golden.py

.PHONY: build test-cpu test-gpu bench golden
