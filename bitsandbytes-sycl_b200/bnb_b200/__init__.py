"""bnb_b200 -- B200 (sm_100a) implementation of the bitsandbytes quantized-linear hot path behind the API
of abhilash1910/bitsandbytes-SYCL's `python_src_quants` package (reference __init__.py:3-11):
`functional`, `matmul`, `matmul_4bit`, `MatmulLtState`, `nn.Linear4bit`, `nn.Linear8bitLt`."""
from . import functional, utils  # noqa: F401
from .autograd._functions import MatmulLtState, matmul, matmul_4bit, matmul_4bit_multi  # noqa: F401
from .nn import modules  # noqa: F401
from . import nn  # noqa: F401

__version__ = "0.43.2.dev+b200.r1"
