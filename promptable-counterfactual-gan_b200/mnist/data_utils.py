"""On-device input pipeline for the MNIST experiments.

Mirror of ``conditional_counteRGAN/mnist/data_utils.py:6-32``:

    get_dataloaders(batch_size=128, num_workers=4, data_dir=None, cfg=None)
        -> (train_loader, valid_loader, test_loader, full_dataset)

The reference decodes ~54 k PIL images per epoch through ``ToTensor`` + ``Normalize((0.5,), (0.5,))`` in DataLoader
workers (:9-12, :26); once the training step takes ~3.5 ms that loader is the bottleneck.  Here the uint8 dataset
(47 MB) stays resident in HBM and a batch is one native kernel (``pcg_u8_batch``): gather through the (shuffled) index
vector + the same normalisation, bit-identical to torchvision's.  The loaders yield ``(x [B,1,28,28] fp32 in [-1,1],
y [B] int64)`` on the device, so ``train_countergan`` consumes them unchanged.  The stratified 90/10 split is the
reference's own ``train_test_split`` call (:19).
"""
import ctypes

import torch

from .. import _lib


def u8_batch(images_u8, labels, index, mean=0.5, std=0.5, out=None):
    """images_u8 [N,H,W] uint8 (CUDA), labels [N] int64 or None, index [B] int64 or None -> (x [B,1,H,W] fp32, y [B])."""
    if not images_u8.is_cuda:
        raise RuntimeError("pcg_b200: the dataset must live on a CUDA device (there is no CPU fallback)")
    N, H, W = images_u8.shape
    B = N if index is None else index.numel()
    x = out if out is not None else torch.empty(B, 1, H, W, device=images_u8.device)
    y = torch.empty(B, dtype=torch.int64, device=images_u8.device) if labels is not None else None
    P = _lib.ptr
    _lib.check(_lib.load().pcg_u8_batch(P(images_u8), P(labels), P(index), B, H * W, ctypes.c_float(mean),
                                        ctypes.c_float(std), P(x), P(y), _lib.stream_ptr()))
    return x, y


class DeviceLoader:
    """Iterable of device batches over a subset of a resident uint8 dataset; ``len`` / ``dataset`` / ``batch_size``
    behave like the DataLoader the reference returns (drop_last=False)."""

    def __init__(self, images_u8, labels, indices=None, batch_size=128, shuffle=False, mean=0.5, std=0.5, generator=None):
        self.images, self.labels = images_u8.contiguous(), labels.contiguous().long()
        dev = self.images.device
        self.indices = (torch.arange(self.images.shape[0], device=dev) if indices is None
                        else torch.as_tensor(indices, dtype=torch.int64, device=dev))
        self.batch_size, self.shuffle, self.mean, self.std, self.generator = batch_size, shuffle, mean, std, generator
        self.dataset = self.indices          # len(loader.dataset) == number of samples, as the reference uses it

    def __len__(self):
        return (self.indices.numel() + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        idx = self.indices
        if self.shuffle:                     # per-epoch permutation drawn on the device
            idx = idx[torch.randperm(idx.numel(), device=idx.device, generator=self.generator)]
        for b in range(len(self)):
            sel = idx[b * self.batch_size:(b + 1) * self.batch_size].contiguous()
            yield u8_batch(self.images, self.labels, sel, self.mean, self.std)


def loaders_from_tensors(train_u8, train_labels, test_u8, test_labels, batch_size=128, device="cuda", seed=None):
    """The body of get_dataloaders for datasets already in memory (uint8 [N,28,28] + labels)."""
    from sklearn.model_selection import train_test_split
    dev = torch.device(device)
    tl = torch.as_tensor(train_labels)
    all_indices = list(range(len(tl)))
    train_idx, valid_idx = train_test_split(all_indices, test_size=0.1, stratify=tl.cpu().numpy(), random_state=seed)
    tr_u8, tr_y = torch.as_tensor(train_u8).to(dev), tl.to(dev)
    te_u8, te_y = torch.as_tensor(test_u8).to(dev), torch.as_tensor(test_labels).to(dev)
    train_loader = DeviceLoader(tr_u8, tr_y, train_idx, batch_size, shuffle=True)
    valid_loader = DeviceLoader(tr_u8, tr_y, valid_idx, batch_size, shuffle=False)
    test_loader = DeviceLoader(te_u8, te_y, None, batch_size, shuffle=False)
    return train_loader, valid_loader, test_loader


def get_dataloaders(batch_size=128, num_workers=4, data_dir=None, cfg=None, device="cuda"):
    """Drop-in for data_utils.py:6-32 (``num_workers`` is accepted and unused: there are no workers)."""
    from torchvision import datasets, transforms
    transform = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,), (0.5,))])
    full_dataset = datasets.MNIST(data_dir, train=True, transform=transform, download=False)   # kept for eval utilities
    test_dataset = datasets.MNIST(data_dir, train=False, transform=transform, download=False)
    train_loader, valid_loader, test_loader = loaders_from_tensors(
        full_dataset.data, full_dataset.targets, test_dataset.data, test_dataset.targets, batch_size, device)
    return train_loader, valid_loader, test_loader, full_dataset
