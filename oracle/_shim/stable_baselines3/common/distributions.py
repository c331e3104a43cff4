class SquashedDiagGaussianDistribution:
    pass
