# Convenience targets (the driver uses __graft_entry__.py / bench.py directly).
PY ?= python

build:            ## nvcc -> pseldnets_b200/libseldfeat.so (sm_100a)
	$(PY) -m pseldnets_b200.build

test-cpu:         ## oracle vs reference goldens, host logic, ABI symbols, gloo sharding
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## parity of the CUDA path (needs a B200)
	$(PY) -m pytest tests -q -m gpu

bench:            ## one JSON line: audio-s/s, roofline, e2e, CPU baseline
	$(PY) bench.py

golden:           ## regenerate tests/golden/*.npz from the unmodified reference (build container only)
	$(PY) tests/golden/make_golden.py
	$(PY) tests/golden/make_golden_epilogue.py
	$(PY) tests/golden/make_golden_augment.py

.PHONY: build test-cpu test-gpu bench golden
