"""Numerical constants of the reference network (reference modules/neural_net/constants.py:9-26)."""
import math

_EPS_ = 1e-5                       # added to the row std in channel_normalization
_LEAKY_RELU_NEG_SLOPE_ = 0.01
_HEAD_WEIGHT_MEAN_INIT_ = 0.0      # final Linear of every task head: N(0, 0.01)
_HEAD_WEIGHT_STD_INIT_ = 0.01
_CLS_BIAS_INIT_ = -math.log(99)    # class heads start at p = 0.01
_REG_BIAS_INIT_ = 0.0
