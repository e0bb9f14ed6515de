"""TEST-ONLY stand-in for `pybullet_data` (reference call site: BaseAviary.py:482)."""


def getDataPath():
    return ""
