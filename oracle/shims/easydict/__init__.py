"""Stand-in for `easydict` (absent in this image). TEST INFRASTRUCTURE ONLY:
lets oracle/ref_harness.py import /root/reference read-only.
Needed by reference isegm/model/is_vpu_model.py:14,190 and isegm/utils/exp.py:10."""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {})
        d.update(kwargs)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v
