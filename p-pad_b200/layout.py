"""Host-side (numpy) description of the packed formats of include/ppad_b200.h.

These helpers build and read the byte layouts the kernels consume -- for file I/O, fixtures and
tests.  They are format converters, not a compute path: no game logic lives here.
"""
import numpy as np


# board sizes with specialised kernels; of those, the ones whose planes are 32-bit words.  Every other size (any
# dim_row x dim_col with at most 64 cells) runs on the run-time-dimension kernels, which use 64-bit planes.
SPECIALISED = ((4, 5), (5, 6), (6, 7))
NARROW = ((4, 5), (5, 6))


def plane_words(rows, cols):
    """uint32 words per bit plane of the packed state"""
    if rows < 1 or cols < 1 or rows * cols > 64:
        raise ValueError("board too large for the packed state (at most 64 cells)")
    return 1 if (rows, cols) in NARROW else 2


def state_words(rows, cols, orb_bits=3):
    """uint32 words per packed state record (orb_bits planes + the meta word, padded to 16 bytes)."""
    if orb_bits not in (3, 4):
        raise ValueError("orb_bits must be 3 or 4")
    return (orb_bits * plane_words(rows, cols) + 1 + 3) // 4 * 4


def pack_state(boards, fingers, prev=None, moves=None, orb_bits=3):
    """[n, rows, cols] ints (-1 = empty, 0..6 or, with orb_bits = 4, 0..9) + [n, 2] (row, col) -> uint32 state records."""
    boards = np.asarray(boards)
    n, rows, cols = boards.shape
    cells = rows * cols
    fingers = np.asarray(fingers).reshape(n, 2)
    codes = boards.reshape(n, cells).astype(np.int64)
    top = 6 if orb_bits == 3 else 9
    if ((codes < -1) | (codes > top)).any():
        raise ValueError("orb types must be -1..%d for the %d-bit state" % (top, orb_bits))
    codes = np.where(codes < 0, (1 << orb_bits) - 1, codes).astype(np.uint64)
    shifts = np.arange(cells, dtype=np.uint64)
    planes = [(((codes >> np.uint64(j)) & np.uint64(1)) << shifts).sum(axis=1, dtype=np.uint64) for j in range(orb_bits)]
    prev = np.full(n, 4, np.uint32) if prev is None else np.asarray(prev).astype(np.uint32)
    moves = np.zeros(n, np.uint32) if moves is None else np.asarray(moves).astype(np.uint32)
    meta = (fingers[:, 0] * cols + fingers[:, 1]).astype(np.uint32) | (prev << np.uint32(6)) | (moves << np.uint32(9))
    out = np.zeros((n, state_words(rows, cols, orb_bits)), np.uint32)
    if plane_words(rows, cols) == 1:
        for j, pl in enumerate(planes):
            out[:, j] = pl.astype(np.uint32)
        out[:, orb_bits] = meta
    else:
        for j, pl in enumerate(planes):
            out[:, 2 * j] = pl.astype(np.uint32)
            out[:, 2 * j + 1] = (pl >> np.uint64(32)).astype(np.uint32)
        out[:, 2 * orb_bits] = meta
    return np.ascontiguousarray(out, np.uint32)


def unpack_state(state, rows, cols, orb_bits=3):
    """uint32 state records -> (boards int8 [n, rows, cols], fingers int32 [n, 2], prev uint8 [n], moves uint16 [n])."""
    state = np.asarray(state, np.uint32)
    n = state.shape[0]
    cells = rows * cols
    if plane_words(rows, cols) == 1:
        planes = [state[:, j].astype(np.uint64) for j in range(orb_bits)]
        meta = state[:, orb_bits]
    else:
        planes = [state[:, 2 * j].astype(np.uint64) | (state[:, 2 * j + 1].astype(np.uint64) << np.uint64(32))
                  for j in range(orb_bits)]
        meta = state[:, 2 * orb_bits]
    shifts = np.arange(cells, dtype=np.uint64)
    codes = sum((((p[:, None] >> shifts) & np.uint64(1)) << np.uint64(j)) for j, p in enumerate(planes)).astype(np.int8)
    boards = np.where(codes == (1 << orb_bits) - 1, -1, codes).astype(np.int8).reshape(n, rows, cols)
    f = (meta & 63).astype(np.int32)
    fingers = np.stack([f // cols, f % cols], 1).astype(np.int32)
    return boards, fingers, ((meta >> 6) & 7).astype(np.uint8), ((meta >> 9) & 0xFFFF).astype(np.uint16)


def pack_slab(orbs):
    """uint8 [n, S] orb sequence -> uint32 [n, ceil(S / 8)] nibble slab (orb d: word d >> 3, bits 4 * (d & 7))."""
    orbs = np.asarray(orbs, np.uint8)
    n, s = orbs.shape
    words = (s + 7) // 8
    pad = np.zeros((n, words * 8), np.uint32)
    pad[:, :s] = orbs & 15
    pad = pad.reshape(n, words, 8)
    return np.ascontiguousarray((pad << (4 * np.arange(8, dtype=np.uint32))).sum(axis=2, dtype=np.uint32))


def pack_slab_planes3(orbs):
    """uint8 [n, S] orb sequence (types 0..6) -> uint32 [n, 3 * ceil(S / 32)] bit-sliced slab (PPAD_SLAB_PLANES3):
    bit (d & 31) of word 3 * (d >> 5) + j is bit j of orb d.  12 bytes per 32 orbs."""
    orbs = np.asarray(orbs, np.uint8)
    n, s = orbs.shape
    groups = max(1, (s + 31) // 32)
    pad = np.zeros((n, groups * 32), np.uint32)
    pad[:, :s] = orbs & 7
    pad = pad.reshape(n, groups, 32)
    shifts = np.arange(32, dtype=np.uint32)
    out = np.zeros((n, groups, 3), np.uint32)
    for j in range(3):
        out[:, :, j] = (((pad >> np.uint32(j)) & np.uint32(1)) << shifts).sum(axis=2, dtype=np.uint32)
    return np.ascontiguousarray(out.reshape(n, groups * 3))


def pack_slab_chunks10(orbs, stride=None):
    """uint8 [n, S] orb sequence (types 0..6) -> uint32 [ceil(S / 10), stride] chunk-major slab (PPAD_SLAB_CHUNKS10):
    bit 10 j + k of out[c, i] is bit j of orb 10 c + k of board i.  4 bytes per 10 orbs; a kernel reading it from pinned
    host memory only fetches the chunks a board's cascade reaches."""
    orbs = np.asarray(orbs, np.uint8)
    n, s = orbs.shape
    chunks = max(1, (s + 9) // 10)
    stride = n if stride is None else int(stride)
    pad = np.zeros((n, chunks * 10), np.uint32)
    pad[:, :s] = orbs & 7
    pad = pad.reshape(n, chunks, 10)
    k = np.arange(10, dtype=np.uint32)
    out = np.zeros((chunks, stride), np.uint32)
    for j in range(3):
        out[:, :n] |= (((pad >> np.uint32(j)) & np.uint32(1)) << (k + np.uint32(10 * j))).sum(axis=2, dtype=np.uint32).T
    return out


def pack_slab_chunks6(orbs, stride=None):
    """uint8 [n, S] orb sequence (colours 0..5) -> uint16 [ceil(S / 6), stride] chunk-major slab (PPAD_SLAB_CHUNKS6):
    out[c, i] = sum_k orb[i, 6 c + k] * 6^k.  2 bytes per 6 orbs."""
    orbs = np.asarray(orbs, np.uint8)
    if orbs.size and orbs.max() > 5:
        raise ValueError("PPAD_SLAB_CHUNKS6 holds the colours 0..5 only")
    n, s = orbs.shape
    chunks = max(1, (s + 5) // 6)
    stride = n if stride is None else int(stride)
    pad = np.zeros((n, chunks * 6), np.uint32)
    pad[:, :s] = orbs
    out = np.zeros((chunks, stride), np.uint16)
    out[:, :n] = (pad.reshape(n, chunks, 6) * (6 ** np.arange(6, dtype=np.uint32))).sum(axis=2).astype(np.uint16).T
    return out


def pack_slab_chunks6t(orbs):
    """uint8 [n, S] orb sequence (colours 0..5) -> uint16 [ceil(n / 256), ceil(S / 6), 256] tile-major slab
    (PPAD_SLAB_CHUNKS6T): out[i // 256, c, i % 256] = sum_k orb[i, 6 c + k] * 6^k."""
    flat = pack_slab_chunks6(orbs)                       # [chunks, n]
    chunks, n = flat.shape
    tiles = max(1, (n + 255) // 256)
    pad = np.zeros((chunks, tiles * 256), np.uint16)
    pad[:, :n] = flat
    return np.ascontiguousarray(pad.reshape(chunks, tiles, 256).transpose(1, 0, 2))


def chunk_bytes_on_demand(info, orbs_per_chunk, chunk_bytes, eager, chunks):
    """Bytes of the chunks beyond the `eager` ones that the cascades of one step reached in a chunk-major slab, from the
    consumed counts in the step's info words (bits 16..23, saturating at 255 orbs): chunk c of a board is read when the
    board drew more than orbs_per_chunk * c orbs; counted in 32-byte sectors of neighbouring boards (a sector is fetched
    once however many of its boards ask for it).  Accounting only: no game logic."""
    sector_boards = 32 // chunk_bytes
    consumed = ((np.asarray(info).view(np.uint32) >> 16) & 0xFF).astype(np.int64)
    consumed = np.pad(consumed, (0, (-len(consumed)) % sector_boards)).reshape(-1, sector_boards).max(axis=1)
    return sum(int((consumed > orbs_per_chunk * c).sum()) * 32 for c in range(eager, chunks))


TILE = 256   # boards per tile of the compact result stream (= CTA size of the step kernel)


def decode_compact(n, words, nibbles, tile_line, records, summary, dictionary=None):
    """Decode the compact result stream of ppad_pipe_step (include/ppad_b200.h: ppad_host_step_args) into dense arrays.
    nibbles uint32 [tiles * 32], tile_line uint32 [tiles], records uint8 [capacity], summary uint32 [4]; words = uint32
    words per packed state record; dictionary = (rewards float64 [D], infos uint32 [D]) for a stream written with
    results_compact = 2.  Returns dict(action uint8 [n], reward float64 [n], info uint32 [n], reset_index int64 [m],
    reset_state uint32 [m, words], lines int, escapes int).  A format decoder: no game logic."""
    if int(summary[2]):
        raise RuntimeError("compact result stream overflowed its record buffer")
    tiles = (n + TILE - 1) // TILE
    nb = np.ascontiguousarray(nibbles[:tiles * 32]).view(np.uint8)
    nib = np.empty(tiles * TILE, np.uint8)
    nib[0::2] = nb & 15
    nib[1::2] = nb >> 4
    nib = nib[:n]
    action = (nib & 7).astype(np.uint8)
    action[action == 7] = 255
    has = (nib >> 3).astype(bool)
    reward = np.zeros(n, np.float64)
    info = np.zeros(n, np.uint32)
    idx = np.flatnonzero(has)
    out = dict(action=action, reward=reward, info=info, lines=int(summary[0]), escapes=0,
               reset_index=np.zeros(0, np.int64), reset_state=np.zeros((0, words), np.uint32))
    if idx.size == 0:
        return out
    tile = idx // TILE
    k = np.bincount(tile, minlength=tiles).astype(np.int64)           # records per tile
    first = np.concatenate([[0], np.cumsum(k)[:-1]])                  # index of each tile's first record in idx
    r = np.arange(idx.size, dtype=np.int64) - first[tile]             # rank within the tile
    base = tile_line[:tiles].astype(np.int64) * 128                   # byte offset of each tile's block
    rec = np.ascontiguousarray(records)
    if dictionary is None:
        reward[idx] = rec.view(np.float64)[base[tile] // 8 + r]
        inf = rec.view(np.uint32)[(base[tile] + 8 * k[tile]) // 4 + r]
        off_state_of = lambda kk, ee: (12 * kk + 15) // 16 * 16
        e = np.zeros(tiles, np.int64)
    else:
        d_reward, d_info = np.asarray(dictionary[0], np.float64), np.asarray(dictionary[1], np.uint32)
        code = rec.view(np.uint16)[base[tile] // 2 + r].astype(np.int64)
        esc = code == 0xFFFF
        if (code[~esc] >= len(d_reward)).any():
            raise RuntimeError("compact result stream: a code is outside the dictionary it is decoded with")
        rw = np.zeros(idx.size, np.float64)
        inf = np.zeros(idx.size, np.uint32)
        rw[~esc] = d_reward[code[~esc]]
        inf[~esc] = d_info[code[~esc]]
        e = np.bincount(tile[esc], minlength=tiles).astype(np.int64)      # escapes per tile
        if esc.any():
            etile = tile[esc]
            efirst = np.concatenate([[0], np.cumsum(e)[:-1]])
            er = np.arange(etile.size, dtype=np.int64) - efirst[etile]
            off_r = base[etile] + (2 * k[etile] + 7) // 8 * 8
            rw[esc] = rec.view(np.float64)[off_r // 8 + er]
            inf[esc] = rec.view(np.uint32)[(off_r + 8 * e[etile]) // 4 + er]
        reward[idx] = rw
        out["escapes"] = int(esc.sum())
        off_state_of = lambda kk, ee: ((2 * kk + 7) // 8 * 8 + 12 * ee + 15) // 16 * 16
    info[idx] = inf
    done = (inf >> 31) != 0
    if done.any():
        didx, dtile = idx[done], tile[done]
        m = np.bincount(dtile, minlength=tiles).astype(np.int64)
        dfirst = np.concatenate([[0], np.cumsum(m)[:-1]])
        rd = np.arange(didx.size, dtype=np.int64) - dfirst[dtile]
        off_state = off_state_of(k[dtile], e[dtile])
        w0 = (base[dtile] + off_state) // 4 + rd * words
        out["reset_index"] = didx
        out["reset_state"] = rec.view(np.uint32)[w0[:, None] + np.arange(words)[None, :]]
    return out
