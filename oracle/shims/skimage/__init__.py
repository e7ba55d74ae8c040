"""Stand-in for `skimage` (absent): measure.label(mask, connectivity=2) via scipy.ndimage
(reference isegm/engine/trainer.py:1176-1177)."""
