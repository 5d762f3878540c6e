"""utils.normalize_image of the reference (utils.py:4-23), on the device.

CUDA tensors are normalised by the grid-patch kernel run with patch size == image size (one patch per image):
bit-exact `(x - min) / ((max - min) + 1e-5)` per image and channel."""
import torch


def normalize_image(image: torch.Tensor) -> torch.Tensor:
    from svrs_native.lib import F32, lib

    if image.ndim not in (3, 4):
        raise ValueError("Input image must be 3D or 4D tensor.")
    if not image.is_cuda:
        raise RuntimeError("normalize_image: svrs_b200 runs on CUDA tensors only (no CPU fallback)")
    img = image.contiguous().float()
    shape = img.shape
    t = img.view(-1, *shape[-3:]) if image.ndim == 4 else img.view(1, *shape)
    n, c, h, w = t.shape
    if h != w:
        raise ValueError("normalize_image kernel expects square images")
    out = torch.empty_like(t)
    lib.grid_patch_normalize(t.data_ptr(), 0, out.data_ptr(), F32, 0, n, c, h, h, torch.cuda.current_stream().cuda_stream)
    return out.view(shape)
