from ...utils import Registry
ATTENTION = Registry('attention')
