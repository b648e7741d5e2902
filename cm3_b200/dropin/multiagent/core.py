"""Host-side attribute records standing in for the reference's entity classes
(multiagent/core.py:4-114).  The trainers and the scenario only read and write attributes on these
objects (`env.world.landmarks[i].state.p_pos`, `agent.size`, `world.dt`, ...); all physics
(World.step, core.py:117-196) runs in the CUDA step kernel, so nothing here computes.

The attribute names and default values are the reference's; they are kept in tables so that the
whole surface is visible at a glance.
"""

# name -> default, per record kind (reference line in the comment)
_STATE_FIELDS = {"p_pos": None, "p_vel": None}                       # core.py:4-10
_AGENT_STATE_FIELDS = dict(_STATE_FIELDS, c=None)                     # core.py:13-17
_ACTION_FIELDS = {"u": None, "c": None}                               # core.py:20-25
_ENTITY_FIELDS = {"name": "", "size": 0.050, "movable": False, "collide": True, "density": 25.0,
                  "color": None, "max_speed": None, "accel": None, "initial_mass": 1.0}  # core.py:28-47
_AGENT_EXTRA = {"movable": True, "silent": False, "blind": False, "u_noise": None, "c_noise": None,
                "u_range": 1.0, "action_callback": None}              # core.py:60-79
_WORLD_FIELDS = {"dim_c": 0, "dim_p": 2, "dim_color": 3,
                 "dt": 0.1, "damping": 0.25, "contact_force": 1e+2, "contact_margin": 1e-3}  # core.py:86-99


class _Record(object):
    """A bag of attributes initialised from a table."""
    _defaults = {}

    def __init__(self):
        for key, value in self._defaults.items():
            setattr(self, key, value)

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ", ".join("%s=%r" % kv for kv in sorted(vars(self).items())))


class EntityState(_Record):
    _defaults = _STATE_FIELDS


class AgentState(_Record):
    _defaults = _AGENT_STATE_FIELDS


class Action(_Record):
    _defaults = _ACTION_FIELDS


class Entity(_Record):
    _defaults = _ENTITY_FIELDS
    _state_cls = EntityState

    def __init__(self):
        _Record.__init__(self)
        self.state = self._state_cls()

    mass = property(lambda self: self.initial_mass)   # core.py:49-51


class Landmark(Entity):
    pass


class Agent(Entity):
    _defaults = dict(_ENTITY_FIELDS, **_AGENT_EXTRA)
    _state_cls = AgentState

    def __init__(self):
        Entity.__init__(self)
        self.action = Action()


class World(_Record):
    _defaults = _WORLD_FIELDS

    def __init__(self):
        _Record.__init__(self)
        self.agents, self.landmarks = [], []

    entities = property(lambda self: self.agents + self.landmarks)                               # core.py:102-104
    policy_agents = property(lambda self: [a for a in self.agents if a.action_callback is None])  # core.py:107-109
    scripted_agents = property(lambda self: [a for a in self.agents if a.action_callback is not None])

    def step(self):
        raise NotImplementedError("World.step is fused into the CUDA step kernel; call "
                                  "MultiAgentEnv.step(action_n)")
