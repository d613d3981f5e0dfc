"""TEST INFRASTRUCTURE ONLY — CPU oracle for the lightprop2d angular-spectrum hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the CPU arm being timed.
The shipped package ``lightprop2d_b200`` never imports this module.
"""
