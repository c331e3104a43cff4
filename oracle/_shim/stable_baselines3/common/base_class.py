class BaseAlgorithm:
    policy = None
    device = "cpu"

    def get_env(self):
        return getattr(self, "env", None)

    def set_env(self, env, force_reset=True):
        self.env = env

    def set_logger(self, logger):
        self._logger = logger

    @property
    def logger(self):
        return self._logger
