"""Host-side mirrors of the reference's entity classes (multiagent/core.py:4-114).  They carry
attributes only - the physics (World.step, core.py:117-196) runs in the CUDA kernel."""
import numpy as np


class EntityState(object):
    def __init__(self):
        self.p_pos = None
        self.p_vel = None


class AgentState(EntityState):
    def __init__(self):
        super(AgentState, self).__init__()
        self.c = None


class Action(object):
    def __init__(self):
        self.u = None
        self.c = None


class Entity(object):
    def __init__(self):
        self.name = ''
        self.size = 0.050
        self.movable = False
        self.collide = True
        self.density = 25.0
        self.color = None
        self.max_speed = None
        self.accel = None
        self.state = EntityState()
        self.initial_mass = 1.0

    @property
    def mass(self):
        return self.initial_mass


class Landmark(Entity):
    def __init__(self):
        super(Landmark, self).__init__()


class Agent(Entity):
    def __init__(self):
        super(Agent, self).__init__()
        self.movable = True
        self.silent = False
        self.blind = False
        self.u_noise = None
        self.c_noise = None
        self.u_range = 1.0
        self.state = AgentState()
        self.action = Action()
        self.action_callback = None


class World(object):
    def __init__(self):
        self.agents = []
        self.landmarks = []
        self.dim_c = 0
        self.dim_p = 2
        self.dim_color = 3
        self.dt = 0.1              # core.py:94
        self.damping = 0.25        # core.py:96
        self.contact_force = 1e+2  # core.py:98
        self.contact_margin = 1e-3  # core.py:99

    @property
    def entities(self):
        return self.agents + self.landmarks

    @property
    def policy_agents(self):
        return [agent for agent in self.agents if agent.action_callback is None]

    @property
    def scripted_agents(self):
        return [agent for agent in self.agents if agent.action_callback is not None]

    def step(self):
        raise NotImplementedError("World.step is fused into the CUDA step kernel; call "
                                  "MultiAgentEnv.step(action_n)")
