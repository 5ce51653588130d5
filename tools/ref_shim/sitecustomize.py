"""Dataset shim for running the UNMODIFIED reference scripts without network access.

Put this directory on PYTHONPATH: ``torchvision.datasets.CIFAR10`` is replaced by a seeded synthetic
dataset with CIFAR-10's exact item contract (PIL-compatible uint8 HxWx3 array → transform, int label;
50 000 train items).  Nothing in the reference source tree is touched."""
import sys


def _install():
    try:
        import numpy as np
        import torchvision
    except Exception:  # torchvision missing: nothing to shim
        return

    class SyntheticCIFAR10:
        def __init__(self, root=None, train=True, transform=None, target_transform=None, download=False):
            n = 50000 if train else 10000
            rng = np.random.default_rng(1234)
            self.targets = rng.integers(0, 10, size=n).tolist()
            protos = rng.integers(0, 256, size=(10, 32, 32, 3), dtype=np.uint8)
            noise = rng.integers(0, 96, size=(n, 32, 32, 3), dtype=np.uint8)
            self.data = (protos[np.asarray(self.targets)] // 2 + noise).astype(np.uint8)
            self.transform, self.target_transform = transform, target_transform
            self.classes = [str(i) for i in range(10)]

        def __len__(self):
            return len(self.targets)

        def __getitem__(self, i):
            img, tgt = self.data[i], self.targets[i]
            if self.transform is not None:
                img = self.transform(img)
            if self.target_transform is not None:
                tgt = self.target_transform(tgt)
            return img, tgt

    torchvision.datasets.CIFAR10 = SyntheticCIFAR10
    torchvision.datasets.cifar.CIFAR10 = SyntheticCIFAR10


if "torch" in sys.modules or True:
    try:
        _install()
    except Exception:
        pass
