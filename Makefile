# Convenience targets; the authoritative entry points are __graft_entry__.py, bench.py and pytest.
PY ?= python

build:            ## nvcc (sm_100a) -> chemsim_b200/libchemsim_lbm.so, C++ driver, CPU oracle
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu:         ## oracle, ABI, host logic, gloo sharding emulation (no GPU)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity through the C ABI on a B200
	$(PY) -m pytest tests -x -q -m gpu

bench:            ## 4096^2 f32 on one GPU, one JSON line
	$(PY) bench.py

bench-cpu:        ## the CPU restatement of lbm.rs on the host cores
	$(PY) bench.py --impl reference

golden:           ## regenerate tests/golden/d2q9_golden.npz from the literal numpy/scipy restatement
	$(PY) tests/golden/make_golden.py

.PHONY: build test-cpu test-gpu bench bench-cpu golden
