"""PAD.__init__ keyword arguments (ppad/pad/game.py:39-50) -> struct ppad_config."""
from . import _cabi
from .pad import parameters


def make_config(dim_row=5, dim_col=6, skyfall=None, skyfall_damage=True, team=None, seed=0,
                cascade_cap=255, reset_cap=64):
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
