"""oracle -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the CUDA hot path.  The product package
(``othello_reinforcement_learning_test_b200``) never imports anything from here;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs do.

* :mod:`oracle.cref`      -- ctypes view of ``ref_rules.c`` (our C restatement)
* :mod:`oracle.refload`   -- loads the compiled reference from ``oracle/_ref``
* :mod:`oracle.net_oracle`-- fp32 eval-mode forward of the reference network
* :mod:`oracle.selfplay_port` -- CPU self-play loop (reference-shaped) for the baseline
"""
