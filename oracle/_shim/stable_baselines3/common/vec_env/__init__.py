"""VecEnv protocol names (step = step_async + step_wait; wrappers pass through)."""


class VecEnv:
    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def step_async(self, actions):
        raise NotImplementedError

    def step_wait(self):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return [None] * self.num_envs


class VecEnvWrapper(VecEnv):
    def __init__(self, venv, observation_space=None, action_space=None):
        self.venv = venv
        VecEnv.__init__(
            self,
            num_envs=venv.num_envs,
            observation_space=observation_space or venv.observation_space,
            action_space=action_space or venv.action_space,
        )

    def step_async(self, actions):
        self.venv.step_async(actions)

    def close(self):
        return self.venv.close()

    def seed(self, seed=None):
        return self.venv.seed(seed)


class DummyVecEnv(VecEnv):
    pass


class SubprocVecEnv(VecEnv):
    pass
