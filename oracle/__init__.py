"""CPU oracle for the self-play hot path of AlphaZero-Piskvorky.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement of the reference's
algorithms (numpy for the board / integer work, torch-CPU functional ops for
the network arithmetic, which in the reference also lives in PyTorch).  It is
the checker the CUDA engine is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it (plus three bring-up diagnostics under
``tools/`` -- ``debug_umma.py``, ``train_diag.py``, ``train_bench.py`` -- which use it
the same way: as the checker, or timed as the CPU baseline of the training step); the product package
(``alphazero-piskvorky_b200/``) never does and has no CPU fallback.

Parity pin: every module here is checked against outputs of the UNMODIFIED
reference (imported from ``/root/reference/alphazero`` in the build container)
that are committed as fixtures under ``tests/golden/`` together with the script
that produced them (``tests/golden/make_golden.py``).  The reference's only own
known-answer test (``alphazero/tests/tests.py:11-22``, the z-labelling truth
table) is restated in ``tests/test_oracle_golden.py``.

All ``file:line`` citations are relative to ``/root/reference/``.
"""

from . import gomoku, mcts, net, selfplay, replay, arena, train  # noqa: F401
