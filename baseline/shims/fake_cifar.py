"""Synthetic stand-in for ``torchvision.datasets.CIFAR10`` (no network, no files): same
interface (``targets``, ``__getitem__`` → (PIL.Image → transform, label)), so the reference's
own transform pipeline and DataLoader run unchanged on CIFAR-10-shaped random images."""
import numpy as np
from PIL import Image


class SyntheticCIFAR10:
    def __init__(self, root=None, train=True, download=False, transform=None, target_transform=None):
        n = 50000 if train else 10000
        rng = np.random.RandomState(0 if train else 1)
        self.targets = [int(i % 10) for i in range(n)]
        self._base = rng.randint(0, 256, size=(256, 32, 32, 3), dtype=np.uint8)   # 256 distinct images, reused
        self.transform, self.target_transform = transform, target_transform

    def __len__(self):
        return len(self.targets)

    def __getitem__(self, i):
        img = Image.fromarray(self._base[i & 255])
        y = self.targets[i]
        if self.transform is not None:
            img = self.transform(img)
        if self.target_transform is not None:
            y = self.target_transform(y)
        return img, y
