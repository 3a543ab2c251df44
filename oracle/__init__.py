"""CPU oracle (test infrastructure only): restatement of nrsyed/pytorch-yolov3's hot path.

Import policy (enforced by tests/test_host_logic.py::test_product_never_imports_the_oracle): only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package.
"""
