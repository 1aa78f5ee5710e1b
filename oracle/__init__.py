"""oracle/ -- CPU restatement of the reference's algorithms for the hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in `gangealing_b200/` imports this package.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline / `--impl reference` legs may import it, and only as the checker or as
the timed CPU baseline -- never as the product.

Pinning: the reference (wpeebles/gangealing) ships no tests and no golden vectors,
so the oracle is pinned against outputs of the reference itself: `oracle/make_golden.py` imports the
read-only reference checkout in the build container (CPU code paths of the reference + torch 2.11),
asserts that every restatement here agrees with it, and writes the seeded input/output fixtures under
`tests/golden/` that travel to the GPU box.  Each function cites the reference file:line it follows.
"""
