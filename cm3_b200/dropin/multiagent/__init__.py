"""Drop-in for the reference's `multiagent` package (env/multiagent-particle-envs/multiagent).
The reference's __init__ only registers gym ids of a non-existent module; nothing to mirror."""
