"""Binds ``nn.Module`` parameters / BatchNorm buffers to the flat arenas the native plan works on.

Works for the mirror classes in ``models/`` and equally for the reference's own modules
(conditional_counteRGAN/mnist/models/*.py): only the ``parameters()`` order and shapes matter, and
those are what ``pcg_mnist_layout`` encodes.
"""
import torch
import torch.nn as nn

from . import plan as P


def _generator_dims(module):
    w = dict(module.named_parameters())
    base_ch = w["conv_in.weight"].shape[0]
    n_res = len(module.resblocks)
    if tuple(w["embed.weight"].shape) != (10, 784) or w["conv_in.weight"].shape[1] != 3:
        raise ValueError("native generator supports img_shape=(1,28,28), num_classes=10 only")
    return base_ch, n_res


def is_bound(module):
    a = getattr(module, "_pcg_arena", None)
    if a is None:
        return False
    return all(p.data_ptr() == ptr for p, ptr in zip(module.parameters(), module._pcg_ptrs))


def bind(module, net, device=None):
    """Moves the module's parameters into a flat CUDA arena (idempotent).  Returns the arena."""
    if is_bound(module):
        return module._pcg_arena
    params = list(module.parameters())
    device = torch.device(device) if device is not None else params[0].device
    if device.type != "cuda":
        raise RuntimeError("pcg_b200: modules must live on a CUDA device (there is no CPU fallback); "
                           "call .to('cuda') first")
    base_ch, n_res = _generator_dims(module) if net == 0 else (64, 6)
    arena = P.Arena(net, base_ch, n_res, device).adopt(params)
    module._pcg_arena = arena
    module._pcg_dims = (base_ch, n_res)
    module._pcg_ptrs = [p.data_ptr() for p in params]
    if net == 0:
        bns = [m for m in module.modules() if isinstance(m, nn.BatchNorm2d)]
        assert len(bns) == 2 * n_res
        running = torch.zeros(2 * n_res, 2, base_ch, dtype=torch.float32, device=device)
        nbt = torch.zeros(2 * n_res, dtype=torch.int64, device=device)
        for i, bn in enumerate(bns):
            running[i, 0].copy_(bn.running_mean)
            running[i, 1].copy_(bn.running_var)
            nbt[i] = int(bn.num_batches_tracked)
            bn._buffers["running_mean"] = running[i, 0]
            bn._buffers["running_var"] = running[i, 1]
            bn._buffers["num_batches_tracked"] = nbt[i]
        module._pcg_bn = (running, nbt)
    module._pcg_versions = None
    return arena


def param_versions(module):
    return tuple(p._version for p in module.parameters())


def dummy_arena(net, device, base_ch=4, n_res=1):
    return P.Arena(net, base_ch, n_res, device)
