"""TEST INFRASTRUCTURE — CPU oracle for the active-perception observation path.

Nothing in the product package (``active_gym_b200``) may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.
"""
