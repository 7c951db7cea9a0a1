"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference hot path.

Nothing under ``oracle/`` is shipped or measured as the product.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker.
"""
