"""Pre-solved boards for smart-data generation (the role of ppad/data/dataset01/solved_boards.py).

The reference ships 16 hand-solved 5x6 boards.  Here solved boards are GENERATED for any board size:
every cell belongs to a full-length or split row/column strip of at least three equal orbs, with
neighbouring strips in different colours, so that a pass on the board cancels every orb.
"""
import numpy as np

# orb types a solved board may contain (solved_boards.py:25)
allowed_orbs = [0, 1, 2, 3, 4, 5]


def _splits(length):
    """ways to cut a line of `length` cells into strips of >= 3 cells"""
    out = [[length]]
    for a in range(3, length - 2):
        out.append([a, length - a])
    return out


def generate_solved_boards(dim_row=5, dim_col=6, count=16, seed=0):
    """`count` solved dim_row x dim_col boards (list of int ndarrays)."""
    rng = np.random.default_rng(seed)
    boards = []
    while len(boards) < count:
        by_rows = bool(rng.integers(2)) if min(dim_row, dim_col) >= 3 else dim_col >= 3
        lines, length = (dim_row, dim_col) if by_rows else (dim_col, dim_row)
        grid = np.zeros((lines, length), dtype=int)
        for i in range(lines):
            cuts = _splits(length)
            cut = cuts[int(rng.integers(len(cuts)))]
            pos = 0
            for seg in cut:
                banned = set()
                if pos > 0:
                    banned.add(int(grid[i, pos - 1]))
                if i > 0:
                    banned.update(int(v) for v in grid[i - 1, pos:pos + seg])
                choices = [c for c in allowed_orbs if c not in banned] or allowed_orbs
                grid[i, pos:pos + seg] = choices[int(rng.integers(len(choices)))]
                pos += seg
        boards.append(grid if by_rows else grid.T.copy())
    return boards


# the default set used by smart_data(): sixteen 5x6 boards, like the reference's fixture
boards = generate_solved_boards(5, 6, 16, seed=2018)
