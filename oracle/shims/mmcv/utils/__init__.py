import logging
import os


class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None):
        self.name = name
        self._module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self._module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self._module_dict.get(key)

    def build(self, cfg, *a, **k):
        return build_from_cfg(cfg, self)


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    t = args.pop('type')
    cls = registry.get(t) if isinstance(t, str) else t
    return cls(**args)


def get_logger(name, log_file=None, log_level=logging.INFO, file_mode='w'):
    return logging.getLogger(name)


def mkdir_or_exist(d, mode=0o777):
    os.makedirs(d, mode=mode, exist_ok=True)
