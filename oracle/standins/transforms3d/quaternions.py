def _na(*a, **k):
    raise NotImplementedError("transforms3d stand-in")


rotate_vector = qconjugate = mat2quat = qmult = _na
