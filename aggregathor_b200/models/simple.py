"""Small models: the MNIST MLP and the CIFAR-10 `cnnet`.

* `mlp`: 784 -> 100 (ReLU) -> 10, variables `dense_i/{weights,biases}` with glorot-uniform weights
  (reference: `experiments/mnist.py:84-104,132`), d = 79 510.
* `cnnet`: conv5x5x3->64 + bias + ReLU + maxpool3/2 SAME; conv5x5x64->64 + bias(0.1) + ReLU + maxpool3/2;
  flatten 8*8*64; dense 4096->384 ReLU; dense 384->192 ReLU; linear 192->10, truncated-normal inits
  (reference: `experiments/cnnet.py:58-95`), d = 1 756 426.
"""

from .core import Conv2d, Dense, Flatten, MaxPool, Model, Sequential


def mlp(dims=(784, 100, 10), name="mlp"):
  layers = [Flatten("flatten")]
  for i in range(len(dims) - 1):
    layers.append(Dense("dense_" + str(i + 1), dims[i], dims[i + 1], relu=(i < len(dims) - 2)))
  return Model(name, Sequential(name, layers), (dims[0],), dims[-1])


def cnnet(num_classes=10, name="cnnet"):
  layers = [
    Conv2d("conv1", 3, 64, 5, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=5e-2, bias_init=0.0),
    MaxPool("pool1", 3, 2, "SAME"),
    Conv2d("conv2", 64, 64, 5, padding="SAME", bias=True, relu=True, init="truncated_normal", init_std=5e-2, bias_init=0.1),
    MaxPool("pool2", 3, 2, "SAME"),
    Flatten("flatten"),
    Dense("dense3", 8 * 8 * 64, 384, relu=True, init="truncated_normal", init_std=0.04, bias_init=0.1),
    Dense("dense4", 384, 192, relu=True, init="truncated_normal", init_std=0.04, bias_init=0.1),
    Dense("linear5", 192, num_classes, relu=False, init="truncated_normal", init_std=1 / 192.0, bias_init=0.0)]
  return Model(name, Sequential(name, layers), (3, 32, 32), num_classes)
