# Convenience targets; the driver uses __graft_entry__.build() / pytest / bench.py directly.
all:
	python -c "import __graft_entry__ as g; g.build()"
test:
	python -m pytest tests -q -m "not gpu"
test-gpu:
	python -m pytest tests -q -m gpu
bench:
	python bench.py
clean:
	$(MAKE) -C nbody-demo-2023_b200 clean
	$(MAKE) -C oracle clean
.PHONY: all test test-gpu bench clean
