"""CPU oracle for the DenseMatcher correspondence hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``densematcher_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` arm do.  See ``oracle/dm_oracle.py`` for the
parity-pinning statement.
"""
