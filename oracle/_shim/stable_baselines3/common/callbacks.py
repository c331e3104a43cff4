class BaseCallback:
    def __init__(self, verbose=0):
        self.verbose = verbose
        self.model = None
        self.logger = None

    def init_callback(self, model):
        self.model = model
        self.logger = model.logger

    def on_rollout_start(self):
        self._on_rollout_start()

    def _on_rollout_start(self):
        pass

    def _on_step(self):
        return True
