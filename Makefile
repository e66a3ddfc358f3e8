# Convenience targets (the reference's Makefile only builds its Sphinx docs).
PY ?= python

.PHONY: build test test-gpu bench smoke clean

build:            ## compile the sm_100a kernels and the C++ helpers in-tree
	$(PY) -c "import __graft_entry__ as g; g.build()"

test:             ## CPU suite (includes multi-process gloo tests)
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## kernel numerics on a B200
	$(PY) -m pytest tests -q -m gpu

smoke:            ## one tiny forward+backward on cuda:0
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench:            ## headline metric on one GPU (use torchrun for N > 1, see bench.py)
	$(PY) bench.py --gpus 1 --steps 3 --warmup 3

clean:
	rm -rf neuronx_distributed_training_b200/ops/_build neuronx_distributed_training_b200/checkpoint/_build build *.egg-info
	find . -name __pycache__ -type d -prune -exec rm -rf {} +
