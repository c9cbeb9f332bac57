# convenience targets (the driver uses __graft_entry__.build())
lib:
	python -c "import lm_b200; lm_b200.build(force=True, verbose=True)"
oracle:
	$(MAKE) -C oracle
cpu-tests:
	python -m pytest tests -x -q -m "not gpu"
.PHONY: lib oracle cpu-tests
