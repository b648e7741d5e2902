class BaseScenario(object):
    """multiagent/scenario.py:4-10"""
    def make_world(self):
        raise NotImplementedError()

    def reset_world(self, world):
        raise NotImplementedError()
