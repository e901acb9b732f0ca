"""CPU oracle for the collaborative-sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU baseline), never as the thing shipped.

Parity status
-------------
* ``sampling_np`` (policy / DRS rejector / MH independence sampler / 2-D
  host-loop refiner): PINNED.  Checked bit-for-bit against the reference's own
  ``sampling/*.py`` executed in the build container (``oracle/ref_shims.py``
  imports them from ``/root/reference`` with three shims); the resulting
  vectors are committed under ``tests/golden/`` by ``oracle/make_golden.py``.
* ``graph_refiner`` + ``nets`` (the TF-1.13 graph refiner over the image
  nets): PARITY UNPINNED.  The reference has no tests, no checkpoints and no
  golden vectors for this arithmetic, and TensorFlow 1.13 cannot be installed
  here; the restatement follows the reference line by line (citations in each
  function) and is cross-checked against ``torch.autograd``.
"""
