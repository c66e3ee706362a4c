"""CPU oracle for the markovflow structured linear-algebra hot path.

TEST INFRASTRUCTURE ONLY. Nothing under ``markovflow_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or as
the timed CPU baseline -- never as the product path.
"""
