import numpy as np
from scipy import ndimage


def label(mask, connectivity=None, background=0, return_num=False):
    nd = np.ndim(mask)
    if connectivity is None:
        connectivity = nd
    st = ndimage.generate_binary_structure(nd, connectivity)
    lab, num = ndimage.label(np.asarray(mask) != background, structure=st)
    return (lab, num) if return_num else lab
