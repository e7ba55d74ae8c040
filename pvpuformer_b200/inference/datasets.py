"""Synthetic evaluation data (the reference's isegm/data package was never published and no dataset is reachable
offline): SURVEY.md 8(d) config 4 -- image i is U[0,1) noise from seed i, its single object a random ellipse."""
import numpy as np


class Sample:
    def __init__(self, image, mask):
        self.image = image
        self._mask = mask
        self.objects_ids = [1]

    def gt_mask(self, object_id):
        return (self._mask == object_id).astype(np.int32)


class SyntheticEllipseDataset:
    def __init__(self, num_images=1024, size=448, seed0=0):
        self.num_images, self.size, self.seed0 = num_images, size, seed0
        self.name = "SyntheticEllipses"
        self._grid = np.mgrid[0:size, 0:size]

    def __len__(self):
        return self.num_images

    def get_sample(self, index):
        rs = np.random.RandomState(self.seed0 + index)
        image = (rs.rand(self.size, self.size, 3) * 255).astype(np.uint8)
        cy, cx = rs.uniform(100, self.size - 100, 2)
        a, b = rs.uniform(40, 140, 2)
        yy, xx = self._grid
        mask = ((((yy - cy) / a) ** 2 + ((xx - cx) / b) ** 2) <= 1).astype(np.int32)
        return Sample(image, mask)


class MaterialisedShard:
    """A rank's block of a dataset generated up front (decoded images in host memory): keeps sample generation out of a timed
    evaluation loop.  Indexable like the dataset itself; only indices of this rank's contiguous shard are held."""

    def __init__(self, ds, rank=0, world=1):
        from .evaluation import shard_range
        self.ds = ds
        self.name = getattr(ds, "name", "dataset")
        a, b = shard_range(len(ds), rank, world)
        self.cache = {i: ds.get_sample(i) for i in range(a, b)}

    def __len__(self):
        return len(self.ds)

    def get_sample(self, index):
        return self.cache[index]
