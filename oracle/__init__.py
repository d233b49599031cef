"""CPU oracle for the fullrmc pair-histogram hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``fullrmc_b200`` never does.

* ``oracle.pairhist``  -- ctypes front-end of ``pairhist_oracle.c`` (C restatement of
  pairs_distances.pyx / pairs_histograms.pyx / reciprocal_space.pyx).
* ``oracle.epilogue``  -- numpy restatement of the constraint-level math
  (__get_total_Gr / __get_total_gr / __get_total_Sq / compute_standard_error).
* ``oracle.build_ref`` -- compiles the REAL reference .pyx files into ``oracle/_ref``
  (``kind: "reference"`` CPU baseline and the authority the restatements are pinned to).

Parity status: pinned (tests/test_oracle.py checks both restatements against
``oracle/_ref`` when it is built and against tests/golden/*.npz always).
"""
