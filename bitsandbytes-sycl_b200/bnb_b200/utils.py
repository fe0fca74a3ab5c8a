"""Host-side helpers mirrored from python_src_quants/utils.py (:169-204): QuantState JSON packing for
state dicts and the Linear8bitLt weight-format map."""
import json

import torch


def pack_dict_to_tensor(source_dict):
    """dict -> uint8 tensor holding its UTF-8 JSON (reference utils.py:169-183)."""
    return torch.tensor(list(json.dumps(source_dict).encode("utf-8")), dtype=torch.uint8)


def unpack_tensor_to_dict(tensor_data):
    """inverse of pack_dict_to_tensor (reference utils.py:186-200)."""
    return json.loads(bytes(tensor_data.cpu().numpy()).decode("utf-8"))


# "row" is what the B200 kernels consume natively; the other three are the reference's IMMA layouts
LINEAR_8BIT_WEIGHTS_FORMAT_MAPPING = {"row": 0, "col32": 1, "col_turing": 2, "col_ampere": 3}
INVERSE_LINEAR_8BIT_WEIGHTS_FORMAT_MAPPING = {v: k for k, v in LINEAR_8BIT_WEIGHTS_FORMAT_MAPPING.items()}
