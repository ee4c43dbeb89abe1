"""Host-side model flattener: pinocchio Model / URDF  ->  KinematicTable  ->  gik_table_t.

The reference builds its model with `RobotWrapper.BuildFromURDF(NEXTAGE_URDF, MESH_DIR)` and then shifts the
robot by editing `model.jointPlacements[1]` (setup_pinocchio.py:28-32, 73-75); the cube's hook frames come from
`cube.data.oMf[getFrameId(LARM_HOOK / RARM_HOOK)]` (setup_pinocchio.py:44-50, tools.py:54-59).  Three
front-ends produce the same POD table:

  * `from_pinocchio(robot, cube)`  -- duck-typed on the pinocchio Model API (used when pinocchio is installed)
  * `from_urdf(robot_urdf, cube_urdf, robot_placement)` -- xml.etree parser, no dependencies
  * `nextage_table()` -- the built-in Nextage + cube_small constants (what the two above return for the
    reference's models; kept so the library works where the URDF files are not available)
"""
from __future__ import annotations

import ctypes
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

GIK_MAX_NQ = 32

# Frame names, config.py:25-29
LEFT_HAND, RIGHT_HAND = "LARM_EFF", "RARM_EFF"
LEFT_HOOK, RIGHT_HOOK = "LARM_HOOK", "RARM_HOOK"


class GikTable(ctypes.Structure):
    """ctypes mirror of gik_table_t (include/gik.h)."""
    _fields_ = [
        ("nq", ctypes.c_int32),
        ("parent", ctypes.c_int32 * GIK_MAX_NQ),
        ("axis", ctypes.c_int32 * GIK_MAX_NQ),
        ("joint_R", (ctypes.c_double * 9) * GIK_MAX_NQ),
        ("joint_p", (ctypes.c_double * 3) * GIK_MAX_NQ),
        ("lower", ctypes.c_double * GIK_MAX_NQ),
        ("upper", ctypes.c_double * GIK_MAX_NQ),
        ("hand_joint", ctypes.c_int32 * 2),
        ("hand_R", (ctypes.c_double * 9) * 2),
        ("hand_p", (ctypes.c_double * 3) * 2),
        ("hook_R", (ctypes.c_double * 9) * 2),
        ("hook_p", (ctypes.c_double * 3) * 2),
    ]


@dataclass
class KinematicTable:
    """Flattened kinematic model, q order = pinocchio joint id - 1 (all joints 1-dof revolute)."""
    names: list
    parent: np.ndarray          # [nq] int, -1 = universe
    axis: np.ndarray            # [nq] int 0/1/2
    joint_R: np.ndarray         # [nq,3,3]
    joint_p: np.ndarray         # [nq,3]
    lower: np.ndarray           # [nq]
    upper: np.ndarray           # [nq]
    hand_joint: np.ndarray      # [2] int
    hand_R: np.ndarray          # [2,3,3]
    hand_p: np.ndarray          # [2,3]
    hook_R: np.ndarray          # [2,3,3]
    hook_p: np.ndarray          # [2,3]
    meta: dict = field(default_factory=dict)

    @property
    def nq(self) -> int:
        return int(len(self.parent))

    def to_c(self) -> GikTable:
        if self.nq > GIK_MAX_NQ:
            raise ValueError(f"nq={self.nq} exceeds GIK_MAX_NQ={GIK_MAX_NQ}")
        t = GikTable()
        t.nq = self.nq
        for i in range(self.nq):
            t.parent[i] = int(self.parent[i])
            t.axis[i] = int(self.axis[i])
            for k in range(9):
                t.joint_R[i][k] = float(self.joint_R[i].reshape(9)[k])
            for k in range(3):
                t.joint_p[i][k] = float(self.joint_p[i][k])
            t.lower[i] = float(self.lower[i])
            t.upper[i] = float(self.upper[i])
        for h in range(2):
            t.hand_joint[h] = int(self.hand_joint[h])
            for k in range(9):
                t.hand_R[h][k] = float(self.hand_R[h].reshape(9)[k])
                t.hook_R[h][k] = float(self.hook_R[h].reshape(9)[k])
            for k in range(3):
                t.hand_p[h][k] = float(self.hand_p[h][k])
                t.hook_p[h][k] = float(self.hook_p[h][k])
        return t

    def to_json(self) -> dict:
        return {
            "names": list(self.names),
            "parent": self.parent.tolist(), "axis": self.axis.tolist(),
            "joint_R": self.joint_R.tolist(), "joint_p": self.joint_p.tolist(),
            "lower": self.lower.tolist(), "upper": self.upper.tolist(),
            "hand_joint": self.hand_joint.tolist(),
            "hand_R": self.hand_R.tolist(), "hand_p": self.hand_p.tolist(),
            "hook_R": self.hook_R.tolist(), "hook_p": self.hook_p.tolist(),
        }

    @staticmethod
    def from_json(d: dict) -> "KinematicTable":
        a = lambda k, dt=np.float64: np.asarray(d[k], dtype=dt)
        return KinematicTable(list(d["names"]), a("parent", np.int64), a("axis", np.int64), a("joint_R"),
                              a("joint_p"), a("lower"), a("upper"), a("hand_joint", np.int64), a("hand_R"),
                              a("hand_p"), a("hook_R"), a("hook_p"))

    def allclose(self, other: "KinematicTable", tol: float = 0.0) -> bool:
        if list(self.names) != list(other.names):
            return False
        for k in ("parent", "axis", "hand_joint"):
            if not np.array_equal(getattr(self, k), getattr(other, k)):
                return False
        for k in ("joint_R", "joint_p", "lower", "upper", "hand_R", "hand_p", "hook_R", "hook_p"):
            if not np.allclose(getattr(self, k), getattr(other, k), rtol=0.0, atol=tol):
                return False
        return True


# ------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------
def rpy_to_matrix(rpy) -> np.ndarray:
    """URDF fixed-axis roll-pitch-yaw: R = Rz(yaw) Ry(pitch) Rx(roll)."""
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _origin(elem):
    o = elem.find("origin")
    if o is None:
        return np.eye(3), np.zeros(3)
    xyz = np.array([float(v) for v in o.get("xyz", "0 0 0").split()])
    rpy = [float(v) for v in o.get("rpy", "0 0 0").split()]
    return rpy_to_matrix(rpy), xyz


def _compose(A, B):
    return A[0] @ B[0], A[1] + A[0] @ B[1]


def _parse_urdf_tree(path):
    """Returns (root_link, children: link -> [joint dict sorted by joint name]).  urdfdom keeps joints in a
    name-sorted map, and pinocchio visits child joints in that order."""
    root = ET.parse(path).getroot()
    joints = []
    child_links = set()
    for j in root.findall("joint"):
        R, p = _origin(j)
        ax = j.find("axis")
        lim = j.find("limit")
        joints.append({
            "name": j.get("name"), "type": j.get("type"),
            "parent": j.find("parent").get("link"), "child": j.find("child").get("link"),
            "R": R, "p": p,
            "axis": None if ax is None else np.array([float(v) for v in ax.get("xyz").split()]),
            "lower": None if lim is None or lim.get("lower") is None else float(lim.get("lower")),
            "upper": None if lim is None or lim.get("upper") is None else float(lim.get("upper")),
        })
        child_links.add(j.find("child").get("link"))
    links = [l.get("name") for l in root.findall("link")]
    roots = [l for l in links if l not in child_links]
    if len(roots) != 1:
        raise ValueError(f"{path}: expected one root link, found {roots}")
    children = {}
    for j in sorted(joints, key=lambda d: d["name"]):
        children.setdefault(j["parent"], []).append(j)
    return roots[0], children


def _flatten_robot(path, root_placement):
    """Depth-first walk merging fixed joints, like pinocchio's URDF parser.  Returns the joint list and a map
    frame name -> (parent q index, (R, p)) for every fixed joint (pinocchio FIXED_JOINT frames)."""
    root_link, children = _parse_urdf_tree(path)
    names, parent, axis, jR, jp, lo, hi = [], [], [], [], [], [], []
    frames = {}

    def visit(link, par_idx, M):  # M = placement of `link` w.r.t. joint par_idx
        for j in children.get(link, []):
            Mj = _compose(M, (j["R"], j["p"]))
            if j["type"] == "fixed":
                frames[j["name"]] = (par_idx, Mj)
                visit(j["child"], par_idx, Mj)
            elif j["type"] in ("revolute", "continuous"):
                a = j["axis"] if j["axis"] is not None else np.array([1.0, 0.0, 0.0])
                k = int(np.argmax(np.abs(a)))
                if not np.allclose(a, np.eye(3)[k]):
                    raise ValueError(f"joint {j['name']}: axis {a} is not a positive coordinate axis")
                idx = len(names)
                names.append(j["name"]); parent.append(par_idx); axis.append(k)
                jR.append(Mj[0]); jp.append(Mj[1])
                lo.append(j["lower"] if j["lower"] is not None else -np.inf)
                hi.append(j["upper"] if j["upper"] is not None else np.inf)
                visit(j["child"], idx, (np.eye(3), np.zeros(3)))
            else:
                raise ValueError(f"joint {j['name']}: unsupported type {j['type']}")

    visit(root_link, -1, (np.eye(3), np.zeros(3)))
    # translaterobot (setup_pinocchio.py:32): jointPlacements[1] = oMf * jointPlacements[1]
    if root_placement is not None:
        Rr, pr = root_placement
        jR[0], jp[0] = _compose((np.asarray(Rr, float), np.asarray(pr, float)), (jR[0], jp[0]))
    return names, parent, axis, jR, jp, lo, hi, frames


def _flatten_cube(path, hooks):
    root_link, children = _parse_urdf_tree(path)
    out = {}

    def visit(link, M):
        for j in children.get(link, []):
            if j["type"] != "fixed":
                raise ValueError("cube model must not have moving joints")
            Mj = _compose(M, (j["R"], j["p"]))
            out[j["name"]] = Mj
            visit(j["child"], Mj)

    visit(root_link, (np.eye(3), np.zeros(3)))
    return [out[h] for h in hooks]


# ------------------------------------------------------------------------------------------------------
# front-ends
# ------------------------------------------------------------------------------------------------------
ROBOT_PLACEMENT = (np.eye(3), np.array([0.0, 0.0, 0.85]))   # config.py:33


def from_urdf(robot_urdf, cube_urdf, robot_placement=ROBOT_PLACEMENT, hands=(LEFT_HAND, RIGHT_HAND),
              hooks=(LEFT_HOOK, RIGHT_HOOK)) -> KinematicTable:
    names, parent, axis, jR, jp, lo, hi, frames = _flatten_robot(robot_urdf, robot_placement)
    for h in hands:
        if h not in frames:
            raise KeyError(f"hand frame {h} not found among the fixed joints of {robot_urdf}")
    hk = _flatten_cube(cube_urdf, hooks)
    return KinematicTable(
        names, np.array(parent, np.int64), np.array(axis, np.int64), np.array(jR), np.array(jp),
        np.array(lo), np.array(hi),
        np.array([frames[h][0] for h in hands], np.int64),
        np.array([frames[h][1][0] for h in hands]), np.array([frames[h][1][1] for h in hands]),
        np.array([m[0] for m in hk]), np.array([m[1] for m in hk]),
        meta={"source": "urdf", "robot_urdf": str(robot_urdf), "cube_urdf": str(cube_urdf)})


_AXIS_OF_SHORTNAME = {"JointModelRX": 0, "JointModelRY": 1, "JointModelRZ": 2}


def from_pinocchio(robot, cube, hands=(LEFT_HAND, RIGHT_HAND), hooks=(LEFT_HOOK, RIGHT_HOOK)) -> KinematicTable:
    """`robot` / `cube`: pinocchio RobotWrapper (or bare Model) as returned by setuppinocchio()
    (setup_pinocchio.py:73-83), i.e. AFTER translaterobot has shifted jointPlacements[1]."""
    model = getattr(robot, "model", robot)
    nj = int(model.njoints)
    names, parent, axis, jR, jp = [], [], [], [], []
    for i in range(1, nj):
        jm = model.joints[i]
        sn = jm.shortname()
        if sn not in _AXIS_OF_SHORTNAME or int(jm.nq) != 1:
            raise ValueError(f"joint {model.names[i]}: unsupported joint model {sn}")
        if int(jm.idx_q) != i - 1:
            raise ValueError("configuration vector is not ordered like the joints")
        names.append(str(model.names[i]))
        parent.append(int(model.parents[i]) - 1)
        axis.append(_AXIS_OF_SHORTNAME[sn])
        M = model.jointPlacements[i]
        jR.append(np.array(M.rotation, float)); jp.append(np.array(M.translation, float).reshape(3))
    lo = np.array(model.lowerPositionLimit, float).reshape(-1)
    hi = np.array(model.upperPositionLimit, float).reshape(-1)

    def frame(m, name):
        fid = m.getFrameId(name)
        if fid >= len(m.frames):
            raise KeyError(f"frame {name} not found")
        f = m.frames[fid]
        pj = f.parentJoint if hasattr(f, "parentJoint") else f.parent
        return int(pj), np.array(f.placement.rotation, float), np.array(f.placement.translation, float).reshape(3)

    hj, hR, hp = zip(*[frame(model, h) for h in hands])
    cmodel = getattr(cube, "model", cube)
    _, kR, kp = zip(*[frame(cmodel, h) for h in hooks])
    return KinematicTable(names, np.array(parent, np.int64), np.array(axis, np.int64), np.array(jR),
                          np.array(jp), lo, hi, np.array([j - 1 for j in hj], np.int64), np.array(hR),
                          np.array(hp), np.array(kR), np.array(kp), meta={"source": "pinocchio"})


def nextage_table() -> KinematicTable:
    """Built-in table for the reference's scene: Nextage (NextageaOpen.urdf:580-730) shifted by
    ROBOT_PLACEMENT (config.py:33, setup_pinocchio.py:32) + cube_small hooks (cube_small.urdf:34-48).
    URDF literals are kept verbatim (1.5708, -3.14; left/right limits differ in the last digits)."""
    names = ["CHEST_JOINT0", "HEAD_JOINT0", "HEAD_JOINT1"] + \
            [f"LARM_JOINT{i}" for i in range(6)] + [f"RARM_JOINT{i}" for i in range(6)]
    parent = [-1, 0, 1, 0, 3, 4, 5, 6, 7, 0, 9, 10, 11, 12, 13]
    ax = {"x": 0, "y": 1, "z": 2}
    arm_axes = [ax[c] for c in "zyyxyz"]
    axis = [ax["z"], ax["z"], ax["y"]] + arm_axes + arm_axes

    def arm(sy):
        return [[0.04, sy * 0.135, 0.1015], [0, 0, 0.066], [0, sy * 0.095, -0.25], [0.1805, 0, -0.03],
                [0.1495, 0, 0], [0, 0, -0.1335]]

    jp = [[0, 0, 0.267 + 0.85], [0, 0, 0.302], [0, 0, 0.08]] + arm(+1.0) + arm(-1.0)
    lower = [-3.14159, -1.22173, -0.401425,
             -1.5707963, -2.44346, -1.22173, -3.1415926, -3.57792, -2.7123889,
             -1.570796, -2.44346, -1.22173, -1.74532, -3.5779, -2.712388]
    upper = [3.14159, 1.22173, 1.308997,
             1.5707963, 1.0471975, 1.5707963, 1.7453292, 1.134464, 2.7123889,
             1.570796, 1.047197, 1.570796, 3.141592, 1.134464, 2.712388]
    eff = rpy_to_matrix([0, 0, 1.5708])
    return KinematicTable(
        names, np.array(parent, np.int64), np.array(axis, np.int64),
        np.tile(np.eye(3), (15, 1, 1)), np.array(jp, float), np.array(lower), np.array(upper),
        np.array([8, 14], np.int64), np.array([eff, eff]),
        np.array([[0.082, 0.05, -0.02], [0.082, -0.05, -0.02]]),
        np.array([np.eye(3), rpy_to_matrix([0, 0, -3.14])]),
        np.array([[0.0, 0.05, 0.0], [0.0, -0.05, 0.0]]),
        meta={"source": "builtin-nextage"})
