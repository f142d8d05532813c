"""CPU oracle for the safeguarded-rollout hot path of TimWalter/SafeGBPO.

TEST INFRASTRUCTURE ONLY.  Nothing under ``safegbpo_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may use it, and there only as the checker or as the timed CPU baseline.

What it is
----------
A plain torch-CPU / numpy / scipy float64 restatement of the reference algorithm:

* ``oracle.dynamics``   simulators, observation maps, rewards, linearisation by
                        ``torch.func.jacrev`` + ``vmap`` exactly as the reference does
                        (src/envs/simulators/interfaces/simulator.py:206-230).
* ``oracle.sets``       zonotope / box helpers, the separating-axis gate
                        (src/sets/box.py:139-154), verdict programs P9 (src/sets/zonotope.py:143-239).
* ``oracle.programs``   the convex programs P1-P4 assembled from the reference's constraint encoders
                        (src/sets/zonotope.py:82-141, src/safeguards/interfaces/safeguard.py:98-208)
                        in their GENERIC lifted form (variables a_s, gamma, Gamma, beta) and solved with
                        scipy-HiGHS / a dense fp64 active-set polish.  No per-environment shortcuts:
                        the CUDA path's closed-form reductions are checked against this.
* ``oracle.safeguard``  Safeguard.actions / BoundaryProjection / RayMask restated
                        (src/safeguards/*.py), with the reference quirks Q1-Q14 of SURVEY.md §0.1.
* ``oracle.shac``       SHAC.collect_trajectories + update_policy (src/learning_algorithms/shac.py:96-151).

Pinning status
--------------
* Simulators, linearisation, reachable set, gate, rollout buffer, SHAC loss/gradients:
  PINNED against the reference's own code, imported unmodified in the build container under
  ``oracle/stubs`` (tests/golden/make_golden.py -> tests/golden/*.npz, checked by
  tests/test_oracle_golden.py).
* Convex programs P1-P4 (the CvxpyLayer solves): **parity unpinned** at the solver boundary.
  cvxpy / cvxpylayers / diffcp / Clarabel are not installable here and the reference's tests
  hold no numeric value for any safe action or gradient (tests/safeguard_test.py:21,37 only count
  non-zero gradient entries).  What pins them instead: the reference's containment verdict
  vectors (tests/sets_test.py:96-101,365-403), the paper's known answers in
  paper_archive/equation_check.ipynb (cells 7, 9, 24, 25: eq. 13-15, 71), scipy-HiGHS as an
  independent LP solver, KKT residual checks and finite differences (tests/test_oracle_programs.py).
"""
