# Convenience targets; the driver uses __graft_entry__.py, pytest and bench.py directly.
PY ?= python

.PHONY: build test-cpu test-gpu smoke bench bench-reference clean

build:            ## libafmg.so (nvcc, sm_100a), the CPU oracle + Hypre stand-in (g++/gcc), the native drivers in tools/
	$(PY) __graft_entry__.py

test-cpu: build   ## oracle, host logic, builders, C / C++ / Fortran-shim consistency (no GPU needed)
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## parity tests through the C ABI (needs a B200)
	$(PY) -m pytest tests -q -m gpu

smoke:
	$(PY) __graft_entry__.py smoke

bench:            ## one JSON line: cell-updates/s on S3, roofline, cpu_baseline, e2e
	$(PY) bench.py

bench-reference:  ## the CPU arm (oracle port on all host cores)
	$(PY) bench.py --impl reference

clean:
	$(MAKE) -C afivo_streamer_b200/csrc clean
	$(MAKE) -C oracle clean
	$(MAKE) -C tools clean
