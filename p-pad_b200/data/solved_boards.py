"""Pre-solved boards for smart-data generation (the role of ppad/data/dataset01/solved_boards.py).

``boards`` are the reference's 16 hand-solved 5x6 boards (solved_boards.py:28-124; the reference writes them 6x5 and
transposes) as a constant table -- data, so that ``ppad.data.dataset01.solved_boards.boards`` and
``smart_data(boards=k)`` draw from the same set as the reference (tests/golden/kat.json pins their combos and rewards).
For other board sizes (BASELINE configs[3]: 7x6) solved boards are GENERATED: every cell belongs to a full-length or
split row/column strip of at least three equal orbs, with neighbouring strips in different colours, so that a pass on
the board cancels every orb.
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


# the reference's sixteen solved 5x6 boards, row-major, one digit per orb (0 red, 1 blue, 2 green, 3 light, 4 dark, 5 heal)
_REFERENCE_BOARDS = [
    "501021 501021 505555 501021 333334",
    "555333 333555 511111 522222 500004",
    "011115 044445 020005 333335 111444",
    "444431 522231 204031 204031 204031",
    "410315 410315 412315 415552 414444",
    "000444 333120 400020 455520 444120",
    "000555 454222 253333 250000 251111",
    "425111 425333 421111 420000 423333",
    "333444 444251 000001 444251 555251",
    "111112 333532 444532 000532 555111",
    "222245 111145 000245 333245 111245",
    "111115 544445 522225 555000 000333",
    "455555 412000 412111 412333 412000",
    "512040 512040 333330 512040 512444",
    "333304 522204 511104 555104 222104",
    "012354 012354 012354 012354 012354",
]
boards = [np.array([[int(ch) for ch in row] for row in text.split()]) for text in _REFERENCE_BOARDS]
