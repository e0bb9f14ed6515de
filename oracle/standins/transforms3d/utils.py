def normalized_vector(*a, **k):
    raise NotImplementedError("transforms3d stand-in")
