from .datasets import CSDataSet, CSDataLoader, DeviceAugment, draw_augmentation, Augmentation  # noqa: F401
