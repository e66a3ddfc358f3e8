# Convenience targets (the reference's Makefile only builds its Sphinx docs).
PY ?= python

.PHONY: build test test-gpu bench smoke clean lint sanitize-host sanitize-gpu sanitize-gpu-small

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

lint:             ## undefined names / unused imports (tools/lint_names.py; no third-party linter needed)
	$(PY) tools/lint_names.py

sanitize-host:    ## TSAN + ASAN/UBSAN builds of the native host code (checkpoint IO threads, dataset index builders)
	$(PY) -m pytest tests/test_native_sanitizers.py -q

# compute-sanitizer over the small-shape kernel tests (run on a B200; every tool replays each kernel, so only the
# small parametrisations are selected).  racecheck covers shared-memory hazards inside a CTA, synccheck the mbarrier /
# bar.sync usage, memcheck out-of-bounds TMA boxes and global accesses.
SANITIZE_K ?= 128-256-64 or 200-264-136 or 72-8-40 or rope or swiglu or sq_norm or 1-256-4-1-0 or grouped_wgrad or bias_act or 64-1024 or embedding
sanitize-gpu:
	@for tool in memcheck racecheck synccheck; do \
	  echo "== compute-sanitizer --tool $$tool"; \
	  timeout 1800 compute-sanitizer --tool $$tool --error-exitcode 1 --target-processes all \
	    $(PY) -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "$(SANITIZE_K)" || exit 1; \
	done

# the quick variant that produced profiles/sanitizer.md: one script with a few CTAs per tcgen05 / TMA kernel
sanitize-gpu-small:
	bash tests/gpu_checks/run_r2_sanitize.sh

clean:
	rm -rf neuronx_distributed_training_b200/ops/_build neuronx_distributed_training_b200/checkpoint/_build build *.egg-info
	find . -name __pycache__ -type d -prune -exec rm -rf {} +
