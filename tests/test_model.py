"""Host-side model flattener (model.py): URDF front-end, pinocchio duck-typed front-end, built-in table and the
committed fixture must all give the same kinematic table.  CPU only."""
import json
import os
import types

import numpy as np
import pytest

from conftest import GOLDEN


def test_builtin_table_equals_fixture_flattened_from_reference_urdfs(table):
    import gik_b200
    fx = gik_b200.KinematicTable.from_json(json.load(open(os.path.join(GOLDEN, "nextage_table.json"))))
    assert table.allclose(fx, 1e-15)
    assert table.nq == 15 and list(table.hand_joint) == [8, 14]
    assert list(table.parent) == [-1, 0, 1, 0, 3, 4, 5, 6, 7, 0, 9, 10, 11, 12, 13]


def test_oracle_constants_typed_independently_agree(table):
    from oracle import grasp_ik_np as o
    assert np.array_equal(table.parent, o.PARENT) and np.array_equal(table.axis, o.AXIS)
    assert np.abs(table.joint_p - o.TRANS).max() < 1e-15
    assert np.array_equal(table.lower, o.LOWER) and np.array_equal(table.upper, o.UPPER)
    for h in (0, 1):
        assert np.abs(table.hand_R[h] - o.HAND_R[h]).max() < 1e-15 and np.abs(table.hook_R[h] - o.HOOK_R[h]).max() < 1e-15


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_from_urdf_merges_fixed_joints_and_sorts_children(tmp_path):
    import gik_b200
    robot = _write(tmp_path, "r.urdf", """<robot name="r">
      <link name="base"/><link name="a"/><link name="b"/><link name="tip"/><link name="mid"/>
      <joint name="Z_second" type="revolute"><parent link="base"/><child link="b"/><origin xyz="0 1 0"/><axis xyz="0 1 0"/><limit lower="-1" upper="2"/></joint>
      <joint name="A_first" type="revolute"><parent link="base"/><child link="mid"/><origin xyz="1 0 0" rpy="0 0 0.5"/><axis xyz="0 0 1"/><limit lower="-3" upper="3"/></joint>
      <joint name="fix" type="fixed"><parent link="mid"/><child link="a"/><origin xyz="0 0 1"/></joint>
      <joint name="HAND_L" type="fixed"><parent link="a"/><child link="tip"/><origin xyz="0.1 0 0" rpy="0 0 1.5708"/></joint>
      <link name="tip2"/><joint name="HAND_R" type="fixed"><parent link="b"/><child link="tip2"/><origin xyz="0 0.2 0"/></joint>
    </robot>""")
    cube = _write(tmp_path, "c.urdf", """<robot name="c"><link name="base_link"/><link name="l"/><link name="r"/>
      <joint name="LARM_HOOK" type="fixed"><parent link="base_link"/><child link="l"/><origin xyz="0 0.05 0"/></joint>
      <joint name="RARM_HOOK" type="fixed"><parent link="base_link"/><child link="r"/><origin xyz="0 -0.05 0" rpy="0 0 -3.14"/></joint></robot>""")
    t = gik_b200.from_urdf(robot, cube, robot_placement=(np.eye(3), [0, 0, 0.85]), hands=("HAND_L", "HAND_R"))
    assert t.names == ["A_first", "Z_second"] and list(t.axis) == [2, 1] and list(t.parent) == [-1, -1]
    assert np.allclose(t.joint_p[0], [1, 0, 0.85])          # root shift applied to the FIRST joint only
    assert np.allclose(t.joint_p[1], [0, 1, 0])
    assert np.allclose(t.hand_p[0], [0.1, 0, 1.0]) and list(t.hand_joint) == [0, 1]
    assert np.allclose(t.hook_R[1], gik_b200.model.rpy_to_matrix([0, 0, -3.14]))


class _SE3:
    def __init__(self, R, p):
        self.rotation, self.translation = np.asarray(R, float), np.asarray(p, float)


def _fake_pinocchio(table):
    """Duck-typed stand-in for pinocchio.Model exposing exactly the attributes from_pinocchio reads."""
    ax = {0: "JointModelRX", 1: "JointModelRY", 2: "JointModelRZ"}
    joints = [types.SimpleNamespace(shortname=lambda: "JointModelRUBX", nq=0, idx_q=-1)]
    for i in range(table.nq):
        joints.append(types.SimpleNamespace(shortname=(lambda a=ax[int(table.axis[i])]: a), nq=1, idx_q=i))
    frames = [types.SimpleNamespace(parentJoint=int(table.hand_joint[h]) + 1, placement=_SE3(table.hand_R[h], table.hand_p[h]))
              for h in (0, 1)]
    names = {"LARM_EFF": 0, "RARM_EFF": 1}
    model = types.SimpleNamespace(
        njoints=table.nq + 1, joints=joints, names=["universe"] + list(table.names),
        parents=[0] + [int(p) + 1 for p in table.parent],
        jointPlacements=[_SE3(np.eye(3), np.zeros(3))] + [_SE3(table.joint_R[i], table.joint_p[i]) for i in range(table.nq)],
        lowerPositionLimit=table.lower, upperPositionLimit=table.upper, frames=frames,
        getFrameId=lambda n: names.get(n, len(frames)))
    cframes = [types.SimpleNamespace(parentJoint=0, placement=_SE3(table.hook_R[h], table.hook_p[h])) for h in (0, 1)]
    cnames = {"LARM_HOOK": 0, "RARM_HOOK": 1}
    cube = types.SimpleNamespace(frames=cframes, getFrameId=lambda n: cnames.get(n, 2))
    return types.SimpleNamespace(model=model), types.SimpleNamespace(model=cube)


def test_from_pinocchio_duck_typed(table):
    import gik_b200
    robot, cube = _fake_pinocchio(table)
    t = gik_b200.from_pinocchio(robot, cube)
    assert t.allclose(table, 0.0)
    with pytest.raises(KeyError):
        gik_b200.from_pinocchio(robot, cube, hands=("NOPE", "RARM_EFF"))


def test_to_c_roundtrip(table):
    c = table.to_c()
    assert c.nq == 15 and c.hand_joint[1] == 14
    assert abs(c.joint_p[0][2] - 1.117) < 1e-15 and c.lower[7] == -3.57792


def _fake_collision_model(scene):
    """Duck-typed pinocchio GeometryModel carrying the packaged reference scene (what scene_from_pinocchio reads)."""
    gos = []
    for g in scene.geoms:
        if g.type == 0:
            geo = types.SimpleNamespace(halfSide=np.array(g.size, float))
        elif g.type == 1:
            geo = types.SimpleNamespace(radius=float(g.size[0]))
        else:
            geo = types.SimpleNamespace(radius=float(g.size[0]), halfLength=float(g.size[1]))
        gos.append(types.SimpleNamespace(name=g.name, geometry=geo, parentJoint=int(g.joint) + 1, placement=_SE3(g.R, g.p)))
    pairs = [types.SimpleNamespace(first=int(a), second=int(b)) for a, b in scene.pairs]
    return types.SimpleNamespace(geometryObjects=gos, collisionPairs=pairs)


def test_scene_from_pinocchio_duck_typed(table):
    # a solver built from a pinocchio RobotWrapper flattens THAT robot's collision model (not a packaged one)
    import gik_b200
    ref = gik_b200.nextage_scene()
    names = {g.name for g in ref.geoms}
    assert "baseLink_0" in names and "obstaclebase_0" in names
    robot, cube = _fake_pinocchio(table)
    robot.collision_model = _fake_collision_model(ref)
    sc = gik_b200.scene_from_pinocchio(robot)
    assert sc.cube == ref.cube and sc.table == ref.table and sc.obstacle == ref.obstacle
    assert np.array_equal(sc.pairs, ref.pairs) and len(sc.geoms) == len(ref.geoms) == 48
    for a, b in zip(sc.geoms, ref.geoms):
        assert a.type == b.type and a.joint == b.joint and np.allclose(a.R, b.R) and np.allclose(a.p, b.p)
        assert np.allclose(np.asarray(a.size)[:len(b.size)], b.size)
    # the cube as hpp-fcl holds it: a mesh (unit cube x meshScale 0.1) -> the box bounding its vertices
    V = np.array([[x, y, z] for x in (-.5, .5) for y in (-.5, .5) for z in (-.5, .5)])
    robot.collision_model.geometryObjects[-1].geometry = types.SimpleNamespace(vertices=lambda: V)
    robot.collision_model.geometryObjects[-1].meshScale = np.array([0.1, 0.1, 0.1])
    sc2 = gik_b200.scene_from_pinocchio(robot)
    assert np.allclose(sc2.geoms[-1].size, [0.05, 0.05, 0.05]) and sc2.geoms[-1].type == 0
    # a moved obstacle is seen (the packaged scene would silently ignore it)
    robot.collision_model.geometryObjects[ref.obstacle].placement = _SE3(np.eye(3), np.array([0.5, 0.2, 0.94]))
    assert np.allclose(gik_b200.scene_from_pinocchio(robot).geoms[ref.obstacle].p, [0.5, 0.2, 0.94])
