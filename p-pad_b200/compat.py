"""install_as_ppad -- register this package under the reference's module names.

After ``ppad_b200.install_as_ppad()``, code written against the reference keeps working unchanged:
``import ppad; ppad.PAD(...)``, ``import ppad.pad.utils as pad_utils``, ``ppad.discount``,
``from ppad.pad.utils import illegal_actions`` (projects/simulation.py:24-27,252;
ppad/data/dataset01/smart_data.py:21-23).  The TensorFlow agent (ppad.agent.agent01.Agent01) is out
of scope and is not provided.
"""
import sys
import types


def install_as_ppad(force=False):
    if "ppad" in sys.modules and not force and not getattr(sys.modules["ppad"], "__ppad_b200__", False):
        raise RuntimeError("a module named 'ppad' is already imported; pass force=True to shadow it")
    import ppad_b200
    from ppad_b200.pad import game, parameters, player, utils as pad_utils
    from ppad_b200.agent import encode, utils as agent_utils
    from ppad_b200.data import smart_data as sd, solved_boards

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__ppad_b200__ = True
        sys.modules[name] = m
        return m

    agent = mod("ppad.agent", utils=agent_utils, encode=encode)
    sys.modules["ppad.agent.utils"] = agent_utils
    pad = mod("ppad.pad", game=game, parameters=parameters, player=player, utils=pad_utils)
    for name, m in (("game", game), ("parameters", parameters), ("player", player), ("utils", pad_utils)):
        sys.modules["ppad.pad." + name] = m
    dataset01 = mod("ppad.data.dataset01", smart_data=sd, solved_boards=solved_boards)
    sys.modules["ppad.data.dataset01.smart_data"] = sd
    sys.modules["ppad.data.dataset01.solved_boards"] = solved_boards
    data = mod("ppad.data", dataset01=dataset01)
    return mod("ppad", PAD=game.PAD, discount=agent_utils.discount, smart_data=sd.smart_data, pad=pad, agent=agent,
               data=data, BatchedPAD=ppad_b200.BatchedPAD, board_finger_to_state=encode.board_finger_to_state)
