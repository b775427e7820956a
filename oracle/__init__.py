"""CPU oracle for the BPR train step + full-catalog evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker / the CPU
baseline, never as the thing shipped.  The product path (``fvx`` package ->
``libfvx.so``) never imports this package and fails loudly without its CUDA
library.

What it restates (all citations relative to the reference tree):

* ``oracle.bpr``       - BPRMF / VBPR forward, loss, closed-form gradients and the
                         Keras-Adam update (``src/recommender/models/BPRMF.py:55-125``,
                         ``src/recommender/models/VBPR.py:59-144``).
* ``oracle.gradfashion`` - the GradFashion linear variant (two-stage visual projection of colour and edge
                         descriptors, ``src/recommender/models/GradFashion.py:81-190,304-320``; SURVEY.md
                         section 8(f) row 4), pinned against the reference's own file; its CUDA path
                         (``FvxModel.two_stage``) is checked against it.
* ``oracle.deferred_adam`` - the lazy form of Keras Adam the CUDA step uses by default (pending step, replayed
                         zero-gradient steps, closed-form decay, flush) next to the literal whole-table form.
* ``oracle.sharded``   - the item-sharded step of ``fvx_bpr_step_sharded`` / ``fvx_bpr_step_sharded_phase``:
                         partial scores, per-rank gradient shares packed by run of equal users, ownership of
                         the per-triple terms - with replicated users and with the users block-owned (run slots
                         by owner; exchanges WU, S, RU, dE); run by tests/test_parallel_cpu.py in two gloo
                         processes.
* ``oracle.tc_bound`` / ``oracle.tc_project`` - the bf16 arithmetic of the two tensor-core paths: operand packing,
                         score bounds, bound encoding and selection rule of the evaluation sweep; hi/lo planes
                         and the three-pass product of the projection.
* ``oracle.sampler``   - the host triple sampler (``src/dataset/dataset.py:83-114``)
                         in the reference's own RNG streams, plus the counter-based
                         Philox sampler the device path implements.
* ``oracle.evaluator`` - candidate lists, AUC/HR/nDCG/P/R and the masked top-k dump
                         (``src/recommender/Evaluator.py:36-128,149-239``).

Pinning status
--------------
The reference ships no tests or golden vectors.  The oracle is pinned against
outputs of the reference's own code run in the build container
(``tests/golden/make_golden.py`` is the generating script; fixtures are committed):

* sampler and evaluator: the reference modules are imported unmodified
  (``dataset.py`` with a stub ``tensorflow`` module, ``Evaluator.py`` as is);
* model math: ``BPRMF.py`` / ``VBPR.py`` are imported unmodified over a small
  torch-backed ``tensorflow`` shim (``tests/golden/tf_shim.py``).  The forward,
  loss and regulariser therefore come from the reference's source; autodiff comes
  from torch; the optimiser arithmetic (``tensorflow==2.3.1`` Keras Adam, not in
  the reference tree and not installable here) is a restatement of its published
  algorithm -> that part is **parity unpinned**.
"""
