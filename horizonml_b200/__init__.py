"""horizonml_b200 — a Blackwell-native hybrid-parallel training framework.

Capabilities mirror HorizonML (three training strategies, a benchmark driver and a
container entrypoint; reference: data_parallel_train.py, layer_model_parallel_train.py,
tensor_parallel_train.py, main.py, train.py) but the design is B200-first:

* one process per GPU, ``torch.distributed`` (NCCL) for bootstrap / p2p only;
* hand-written sm_100a kernels (``csrc/``): tcgen05/TMEM/TMA implicit-GEMM
  convolutions, fused BatchNorm/ReLU/residual, fused FC+softmax-CE head, fused Adam,
  peer-memory one-shot / two-shot / NVLS all-reduce with cast+scale fused in,
  GEMM+reduce-scatter and all-gather+GEMM for tensor parallelism;
* CUDA-graph captured training step, device-side metrics (no host sync per step);
* a pure-PyTorch op backend (``ops.backend('torch')``) that doubles as the numerical
  oracle and as the CPU/gloo plumbing path used by the test-suite.
"""

__version__ = "0.1.0"

from .config import TrainConfig  # noqa: F401
