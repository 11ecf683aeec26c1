"""CPU oracle for the SS2D hot path of MedMamba.

TEST INFRASTRUCTURE ONLY.  Nothing under ``medmamba_b200/`` imports this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import, call, link or execute anything under
``oracle/``, and there only as the checker or the timed CPU baseline -- never as
the product path.

Parity status
-------------
* The scan arithmetic (``selective_scan_ref``) lives in the un-vendored third-party
  package ``mamba_ssm`` (README.md:19 pins ``mamba_ssm==1.0.1``); the reference tree
  holds only its text, pasted as docstrings at ``temp.py:57-139``.  The reference
  ships no tests, golden vectors or fixtures for this boundary, so by the
  reference's own material the scan boundary is **parity unpinned**.  We pin the
  restatement in ``oracle/selective_scan_ref.py`` ourselves: ``oracle/refload.py``
  (container only) *executes the docstring text of temp.py:57-139 itself* and
  ``oracle/make_golden.py`` stores its outputs under ``tests/golden/``.
* Everything else on the path (cross-scan, x_proj/dt_proj, cross-merge, out_norm,
  gate, channel_shuffle, the VSSM backbone) is the reference's own torch code in
  ``MedMamba.py``; ``oracle/refload.py`` imports that file UNMODIFIED from
  ``/root/reference`` (container only) and the golden fixtures hold its outputs.
"""
