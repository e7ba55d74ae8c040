class FileClient:
    def __init__(self, *a, **k):
        raise NotImplementedError


def load(*a, **k):
    raise NotImplementedError
