from ._functions import MatMul4Bit, MatMul8bitLt, MatmulLtState, matmul, matmul_4bit  # noqa: F401
