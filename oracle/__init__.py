"""CPU oracle for the ApplyMasksUDF / CoM hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy + a little C) of the reference
LiberTEM algorithm for the masked-reduction hot path.  It exists to *check*
the CUDA product path; it is never imported by ``libertem_b200`` itself.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.

Parity status: PINNED.  Every function here is checked against outputs of the
unmodified reference (run in the build container from /root/reference/src under
third-party shims, see tests/golden/make_golden.py) that are committed under
tests/golden/*.npz; see tests/test_oracle_golden.py.

Each function cites the reference file:line it restates (paths relative to
/root/reference/).
"""
