"""TEST INFRASTRUCTURE ONLY: CPU oracle for the CBAM / SwinBlock / SPPF hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this package.  The product package never does (tests/test_boundary.py greps for that).
"""
