"""TEST-ONLY stand-in for the `pybullet` C extension (absent from this image).

ORACLE / TEST INFRASTRUCTURE -- never imported by the product package.

It lets the UNMODIFIED reference sources under /root/reference run their
`Physics.DYN` path, which uses Bullet only as (a) a per-body state store and
(b) three pure quaternion helpers.  Call sites in the reference:
  state store   gym_pybullet_drones/envs/BaseAviary.py:479-491, 517-519, 865-875
  helpers       gym_pybullet_drones/envs/BaseAviary.py:488,518,836
                gym_pybullet_drones/control/DSLPIDControl.py:144,187,240,241
The helper formulas restate Bullet's published algorithms (pybullet ^3.2.7,
pyproject.toml:19; btMatrix3x3::setRotation, pybullet.c getEulerFromQuaternion /
getQuaternionFromEuler).  The library itself is not available, so parity at
this boundary is cross-checked against scipy only ("parity unpinned" vs Bullet).

`applyExternalForce/Torque` do not integrate anything: they append to
`APPLIED` so tests can pin the reference's _groundEffect/_drag/_downwash
force formulas (BaseAviary.py:715-811) at formula level.  `stepSimulation`
raises: Bullet's solver (Physics.PYB*) is out of scope.
"""
import math
import os
import xml.etree.ElementTree as _et

GUI = 1
DIRECT = 2
LINK_FRAME = 1
WORLD_FRAME = 2
URDF_USE_INERTIA_FROM_FILE = 2
COV_ENABLE_RGB_BUFFER_PREVIEW = 0
COV_ENABLE_DEPTH_BUFFER_PREVIEW = 1
COV_ENABLE_SEGMENTATION_MARK_PREVIEW = 2
ER_TINY_RENDERER = 0
ER_SEGMENTATION_MASK_OBJECT_AND_LINKINDEX = 1
STATE_LOGGING_VIDEO_MP4 = 3

_BODIES = {}
_NEXT_ID = [0]
APPLIED = []          # (kind, body, link, vec3, flags)


def connect(mode, *a, **k):
    return 0


def disconnect(*a, **k):
    pass


def resetSimulation(*a, **k):
    _BODIES.clear()
    _NEXT_ID[0] = 0
    APPLIED.clear()


def setGravity(*a, **k):
    pass


def setRealTimeSimulation(*a, **k):
    pass


def setTimeStep(*a, **k):
    pass


def setAdditionalSearchPath(*a, **k):
    pass


def configureDebugVisualizer(*a, **k):
    pass


def _link_offsets(path):
    """Inertial-origin xyz of every non-base link, in URDF order."""
    offs = []
    if not os.path.isfile(path):
        return offs
    root = _et.parse(path).getroot()
    links = [c for c in root if c.tag == 'link']
    for ln in links[1:]:
        o = ln.find('inertial/origin')
        offs.append([float(s) for s in o.attrib['xyz'].split()] if o is not None else [0., 0., 0.])
    return offs


def loadURDF(fileName, basePosition=(0., 0., 0.), baseOrientation=(0., 0., 0., 1.), *a, **k):
    bid = _NEXT_ID[0]
    _NEXT_ID[0] += 1
    _BODIES[bid] = dict(pos=tuple(float(v) for v in basePosition),
                        quat=tuple(float(v) for v in baseOrientation),
                        vel=(0., 0., 0.), ang=(0., 0., 0.),
                        links=_link_offsets(fileName))
    return bid


def resetBasePositionAndOrientation(bodyUniqueId, posObj, ornObj, *a, **k):
    b = _BODIES[int(bodyUniqueId)]
    b['pos'] = tuple(float(v) for v in posObj)
    b['quat'] = tuple(float(v) for v in ornObj)


def resetBaseVelocity(objectUniqueId, linearVelocity=None, angularVelocity=None, *a, **k):
    b = _BODIES[int(objectUniqueId)]
    if linearVelocity is not None:
        b['vel'] = tuple(float(v) for v in linearVelocity)
    if angularVelocity is not None:
        b['ang'] = tuple(float(v) for v in angularVelocity)


def getBasePositionAndOrientation(bodyUniqueId, *a, **k):
    b = _BODIES[int(bodyUniqueId)]
    return b['pos'], b['quat']


def getBaseVelocity(bodyUniqueId, *a, **k):
    b = _BODIES[int(bodyUniqueId)]
    return b['vel'], b['ang']


def getMatrixFromQuaternion(q, *a, **k):
    x, y, z, w = (float(v) for v in q)
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return (1.0 - (yy + zz), xy - wz, xz + wy,
            xy + wz, 1.0 - (xx + zz), yz - wx,
            xz - wy, yz + wx, 1.0 - (xx + yy))


def getEulerFromQuaternion(q, *a, **k):
    x, y, z, w = (float(v) for v in q)
    sqx, sqy, sqz, squ = x * x, y * y, z * z, w * w
    sarg = -2.0 * (x * z - w * y)
    if sarg <= -0.99999:
        return (0.0, -0.5 * math.pi, 2.0 * math.atan2(x, -y))
    if sarg >= 0.99999:
        return (0.0, 0.5 * math.pi, 2.0 * math.atan2(-x, y))
    return (math.atan2(2.0 * (y * z + w * x), squ - sqx - sqy + sqz),
            math.asin(sarg),
            math.atan2(2.0 * (x * y + w * z), squ + sqx - sqy - sqz))


def getQuaternionFromEuler(e, *a, **k):
    r, p_, y_ = (float(v) for v in e)
    cr, sr = math.cos(r * 0.5), math.sin(r * 0.5)
    cp, sp = math.cos(p_ * 0.5), math.sin(p_ * 0.5)
    cy, sy = math.cos(y_ * 0.5), math.sin(y_ * 0.5)
    x = sr * cp * cy - cr * sp * sy
    y = cr * sp * cy + sr * cp * sy
    z = cr * cp * sy - sr * sp * cy
    w = cr * cp * cy + sr * sp * sy
    n = math.sqrt(x * x + y * y + z * z + w * w)
    return (x / n, y / n, z / n, w / n)


def getLinkStates(bodyUniqueId, linkIndices, *a, **k):
    """[0] of each entry = world position of the link's COM (what _groundEffect reads)."""
    b = _BODIES[int(bodyUniqueId)]
    R = getMatrixFromQuaternion(b['quat'])
    out = []
    for li in linkIndices:
        o = b['links'][li]
        wp = tuple(b['pos'][r] + R[3 * r] * o[0] + R[3 * r + 1] * o[1] + R[3 * r + 2] * o[2] for r in range(3))
        out.append((wp, b['quat'], (0., 0., 0.), (0., 0., 0., 1.), wp, b['quat'], b['vel'], b['ang']))
    return out


def applyExternalForce(objectUniqueId, linkIndex, forceObj, posObj, flags, *a, **k):
    APPLIED.append(('force', int(objectUniqueId), int(linkIndex), tuple(float(v) for v in forceObj), flags))


def applyExternalTorque(objectUniqueId, linkIndex, torqueObj, flags, *a, **k):
    APPLIED.append(('torque', int(objectUniqueId), int(linkIndex), tuple(float(v) for v in torqueObj), flags))


def stepSimulation(*a, **k):
    raise NotImplementedError("stand-in pybullet: Bullet's rigid-body solver (Physics.PYB*) is not available")
