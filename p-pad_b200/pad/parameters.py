"""Game constants of the reference environment (ppad/pad/parameters.py): orb-type names, default
skyfall rates, default team, match length.  Plain data; the kernels receive it through ppad_config."""

# ppad/pad/parameters.py:20-29
int2english = {0: 'red', 1: 'blue', 2: 'green', 3: 'light', 4: 'dark', 5: 'heal',
               6: 'jammer', 7: 'poison', 8: 'mortal poison', 9: 'bomb'}
english2int = {name: i for i, name in int2english.items()}

# ppad/pad/parameters.py:43-51: rates for the ten orb types, in int2english order
default_skyfall = [1/6, 1/6, 1/6, 1/6, 1/6, 1/6, 0, 0, 0, 0]
default_skyfall_locked = [0] * 10
default_skyfall_enhanced = [0] * 10


def _unit(name, c1, c2):
    return {'name': name, 'color_1': c1, 'color_2': c2, 'atk': 1000, 'hp': 2000, 'rcv': 100}


# ppad/pad/parameters.py:54-89: two leaders and four subs
default_team = [_unit('Leader 1', 'dark', 'dark'), _unit('Leader 2', 'light', 'light'),
                _unit('Sub 1', 'blue', 'green'), _unit('Sub 2', 'green', 'red'),
                _unit('Sub 3', 'red', 'blue'), _unit('Sub 4', 'blue', None)]

default_enemy = []            # parameters.py:93
default_finger = [0, 0]       # parameters.py:96

# ppad/pad/parameters.py:101-111: (letter, ANSI colour) per orb type for terminal rendering
default_render_dict = {-1: (' ', ''),
                       0: ('R', '\033[1;31;41m'), 1: ('B', '\033[1;34;44m'), 2: ('G', '\033[1;32;42m'),
                       3: ('L', '\033[1;33;43m'), 4: ('D', '\033[1;35;45m'), 5: ('H', '\033[1;30;40m'),
                       6: ('j', '\033[1;30;16m'), 7: ('p', '\033[1;5;16m'), 8: ('m', '\033[1;90;16m'),
                       9: ('b', '\033[1;94;16m')}

# parameters.py:114 -- the reference's windows hard-code 3 (utils.py:49-54); only 3 is supported
match_n = 3
