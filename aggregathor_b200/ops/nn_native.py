"""sm_100a providers of the nn ops (`native/op_nn`). Every function returns None (or NotImplemented for
backward ops that return None legitimately) when it does not handle the given arguments, in which case
`ops/nn.py` decides between falling back and failing (AGB_NATIVE_STRICT)."""


def conv2d_forward(x, weight, bias, stride, pads, relu):
  return None


def conv2d_backward(dy, x, weight, y, stride, pads, relu, has_bias, need_dx, grad_w, grad_b):
  return NotImplemented


def linear_forward(x, weight, bias, relu):
  return None


def linear_backward(dy, x, weight, y, relu, need_dx, grad_w, grad_b):
  return NotImplemented


def batchnorm_forward(x, gamma, beta, moving_mean, moving_var, decay, eps, relu):
  return None


def batchnorm_backward(dy, x, y, gamma, mean, rstd, relu, grad_gamma, grad_beta):
  return None


def relu_backward(dy, y):
  return None


def add_relu_forward(a, b, relu):
  return None


def maxpool_forward(x, k, stride, pads):
  return None


def maxpool_backward(dy, x, y, k, stride, pads):
  raise NotImplementedError


def global_avgpool_forward(x):
  return None


def global_avgpool_backward(dy, shape):
  return None


def softmax_xent(logits, labels, label_smoothing):
  return None


def image_normalize(images, mode, dtype):
  return None
