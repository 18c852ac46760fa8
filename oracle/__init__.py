"""TEST INFRASTRUCTURE ONLY — the CPU oracle for the blob-splat hot path.

Nothing under ``oracle/`` is part of the product. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  The product
package ``blobctrl_b200`` never imports this package and has no CPU fallback.

Parity pinning: the reference (TencentARC/BlobCtrl) ships no tests or golden vectors for
this path.  The oracle is therefore pinned against outputs of the
reference itself, executed in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference/blobctrl/utils/utils.py`` unmodified) and committed as
fixtures under ``tests/golden/``.
"""
