"""PAD.__init__ keyword arguments (ppad/pad/game.py:39-50) -> struct ppad_config."""
from . import _cabi
from .pad import parameters


def make_config(dim_row=5, dim_col=6, skyfall=None, skyfall_damage=True, team=None, seed=0,
                cascade_cap=255, reset_cap=64, orb_bits=None):
    """orb_bits: 3 (orb types 0..6), 4 (all ten types of game.py:236-246) or None = 4 exactly when the skyfall
    produces one of the types 7..9 (poison, mortal poison, bomb)."""
    cfg = _cabi.default_config()
    cfg.rows, cfg.cols = int(dim_row), int(dim_col)
    cfg.skyfall_damage = int(bool(skyfall_damage))
    cfg.cascade_cap, cfg.reset_cap = int(cascade_cap), int(reset_cap)
    cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    if skyfall is not None:
        if len(skyfall) != 10:
            raise Exception('ERROR: skyfall needs the rates of the 10 orb types!')
        for i, p in enumerate(skyfall):
            cfg.skyfall[i] = float(p)
    if orb_bits is None:
        orb_bits = 4 if skyfall is not None and any(float(p) > 0 for p in skyfall[7:]) else 3
    if orb_bits not in (3, 4):
        raise ValueError("orb_bits must be 3 or 4")
    cfg.orb_bits = int(orb_bits)
    if team is not None:
        if len(team) > _cabi.MAX_TEAM_UNITS:
            raise Exception('ERROR: at most %d team units are supported' % _cabi.MAX_TEAM_UNITS)
        cfg.team_size = len(team)
        for i, unit in enumerate(team):
            for field, name in ((cfg.team_color_1, 'color_1'), (cfg.team_color_2, 'color_2')):
                colour = unit[name]
                field[i] = -1 if colour is None else parameters.english2int[colour]
            cfg.team_atk[i] = float(unit['atk'])
    return cfg
