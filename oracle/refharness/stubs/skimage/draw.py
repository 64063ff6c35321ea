def circle(*a, **k):
    raise NotImplementedError
