"""CPU oracle for the bayes-kit sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``bayes-kit_b200/``) imports this directory.  The only legal callers are
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker
or as the timed CPU baseline, never as the thing shipped.

The oracle is a plain NumPy restatement of the algorithms in
flatironinstitute/bayes-kit (pure Python + NumPy upstream), one function per
reference entry point, each citing the reference ``file:line`` it follows.
It is *pinned*: ``oracle/gen_golden.py`` runs the unmodified reference (from
``/root/reference``) under a recording RNG proxy, checks that the restatement
reproduces it, and writes the recorded streams + reference outputs to
``tests/golden/*.npz``; ``tests/test_oracle_*.py`` re-check the restatement
against those fixtures on every run (and against the live reference whenever
it is importable).
"""
