"""Shared test helpers: torch <-> oracle bit-pattern bridges and error metrics."""
import numpy as np
import torch

DT = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}


def to_bits(t: torch.Tensor) -> np.ndarray:
    """torch tensor -> numpy array the oracle understands (fp32 as float32, 16-bit as uint16 bit patterns)."""
    t = t.detach().contiguous().cpu()
    if t.dtype == torch.float32:
        return t.numpy()
    if t.dtype in (torch.float16, torch.bfloat16):
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def from_bits(a: np.ndarray, dtype: str) -> torch.Tensor:
    if dtype == "fp32":
        return torch.from_numpy(np.ascontiguousarray(a, np.float32))
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(DT[dtype])


def bits_equal(t: torch.Tensor, a: np.ndarray) -> bool:
    return np.array_equal(to_bits(t).ravel().view(np.uint8), np.ascontiguousarray(a).ravel().view(np.uint8))


def rel_l2(y, ref) -> float:
    y = np.asarray(y, np.float64).ravel()
    ref = np.asarray(ref, np.float64).ravel()
    return float(np.linalg.norm(y - ref) / max(np.linalg.norm(ref), 1e-30))


def adversarial_block_values(n_blocks: int, blocksize: int, thresholds, seed: int = 0) -> np.ndarray:
    """fp32 data whose normalised values sit on / next to every decision threshold, plus zero blocks,
    denormals, +-absmax ties and huge / tiny scales."""
    rng = np.random.RandomState(seed)
    A = rng.randn(n_blocks, blocksize).astype(np.float32)
    thr = np.asarray(thresholds, np.float32)
    for b in range(0, n_blocks, 3):
        scale = np.float32(2.0 ** rng.randint(-20, 20))
        A[b, 0] = scale                       # absmax exactly a power of two: x / absmax is exact
        A[b, 1] = -scale
        k = min(len(thr), (blocksize - 2) // 3)
        A[b, 2:2 + k] = thr[:k] * scale
        A[b, 2 + k:2 + 2 * k] = np.nextafter(thr[:k], np.float32(2)) * scale
        A[b, 2 + 2 * k:2 + 3 * k] = np.nextafter(thr[:k], np.float32(-2)) * scale
    if n_blocks > 4:
        A[1] = 0.0                            # all-zero block: absmax 0 -> inv inf -> NaN -> code 0
        A[4, :] = np.float32(1e-42)           # denormals
        A[4, 3] = np.float32(-3e-41)
    return A
