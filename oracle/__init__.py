"""ORACLE package — test infrastructure only.

CPU restatement of the rasterization hot path of inuex35/splat_one's gsplat fork.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it; `splat_one_b200/` never does.  See oracle/torch_ref.py
for the parity status of each function.
"""
