"""TEST INFRASTRUCTURE ONLY -- CPU/fp32 restatement of the ProPainter inference hot path.

Nothing in ``propainter_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference``
legs use it, and only as the checker / the baseline, never as the product.

Every function is a plain-PyTorch (fp32) restatement of one reference function and
cites the reference file:line it follows.  The functions are *functional*: they
take a flat ``state_dict`` (reference key names) instead of ``nn.Module`` objects,
so the same weights drive the oracle, the reference modules (in the authoring
container, see ``tests/golden/make_golden.py``) and the CUDA product.

Parity pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md §4, §8c).  The oracle is pinned against outputs of the *reference's own
modules* run in the authoring container on seeded inputs; those outputs are
committed under ``tests/golden/`` together with the generating script.
"""
