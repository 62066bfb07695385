"""TEST INFRASTRUCTURE ONLY — CPU oracle for the NoiseDiff reverse-diffusion hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker.
The shipped path (``noisediff_b200``) never imports this package and fails loudly without its CUDA library.
"""
