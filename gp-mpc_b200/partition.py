"""How the independent per-output GPs are split over GPU ranks (index work: exact).

The reference loops ``for output in range(Ny)`` (optimize.py:433,
gp_functions.py:128); those iterations are independent, so rank r owns the
contiguous block ``[r*ceil(Ny/W), ...)``.  When that would leave a rank empty
(Ny < W or ragged tail) the model is replicated instead and the H test points
of a batch are split (second partitioning of SURVEY.md section 8e).
"""
from __future__ import annotations


def output_block(Ny: int, rank: int, world: int):
    """(begin, count) of the outputs owned by `rank`; count may be 0."""
    if world < 1 or not (0 <= rank < world) or Ny < 1:
        raise ValueError('bad partition request Ny=%d rank=%d world=%d' % (Ny, rank, world))
    per = (Ny + world - 1) // world
    begin = min(Ny, rank * per)
    return begin, max(0, min(per, Ny - begin))


def choose_mode(Ny: int, world: int) -> str:
    """'outputs' when every rank owns at least one output, else 'points'."""
    if world == 1:
        return 'outputs'
    return 'outputs' if all(output_block(Ny, r, world)[1] > 0 for r in range(world)) else 'points'


def point_block(H: int, rank: int, world: int):
    """(begin, count) of the test points rank handles in 'points' mode."""
    per = (H + world - 1) // world
    begin = min(H, rank * per)
    return begin, max(0, min(per, H - begin))
