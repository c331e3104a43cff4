def check_for_correct_spaces(env, observation_space, action_space):
    if observation_space != env.observation_space:
        raise ValueError("Observation spaces do not match")
    if action_space != env.action_space:
        raise ValueError("Action spaces do not match")
