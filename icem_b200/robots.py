"""Articulated-body model descriptions for the ground-truth forward models and their flat tables.

The reference's ground-truth dynamics is MuJoCo 2.0 driven through gym's `half_cheetah.xml` /
`humanoidstandup.xml` (icem/environments/mujoco.py:2-10, :48-131, :228-277); neither MuJoCo nor the XML files are
available here, so the kinematic trees, joint ranges / stiffness / damping / armature, actuator gears
and capsule geometry below are re-typed from memory of those public model files and simulated
by this repo's own rigid-body engine (csrc/dyn_articulated.cuh).  PARITY WITH MUJOCO IS UNPINNED; what is pinned is
the CUDA engine against the independent float64 restatement in oracle/articulated_np.py on these same tables.

`compile_model()` turns a description into the flat float32/int32 tables the C ABI takes
(`icem_articulated_model_t`, include/icem_b200.h).  This module is DATA + geometry preprocessing only (mass and
inertia of capsules / spheres, contact points); it contains no dynamics.
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

SLIDE, HINGE, FREE_TRANS, FREE_ROT = 0, 1, 2, 3
MAX_BODIES, MAX_DOFS, MAX_CONTACTS, MAX_CHILDREN = 16, 32, 32, 4


@dataclass
class Joint:
    name: str
    kind: str                     # "slide" | "hinge" | "free"
    axis: Tuple[float, float, float] = (0.0, 0.0, 1.0)
    pos: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    range: Optional[Tuple[float, float]] = None     # radians / metres; None = unlimited
    stiffness: float = 0.0
    damping: float = 0.0
    armature: float = 0.0


@dataclass
class Geom:
    name: str
    kind: str                     # "capsule" | "sphere"
    size: float                   # radius
    fromto: Optional[Tuple[float, ...]] = None      # capsule end points (body frame)
    pos: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    half: float = 0.0             # capsule half length (with `yangle`)
    yangle: float = 0.0           # capsule axis = R_y(yangle) e_z  (MuJoCo axisangle="0 1 0 yangle")

    def endpoints(self):
        if self.kind == "sphere":
            return [np.asarray(self.pos, float)]
        if self.fromto is not None:
            f = np.asarray(self.fromto, float)
            return [f[:3], f[3:]]
        d = np.array([math.sin(self.yangle), 0.0, math.cos(self.yangle)])
        c = np.asarray(self.pos, float)
        return [c - self.half * d, c + self.half * d]


@dataclass
class Body:
    name: str
    parent: int                   # index into the body list, -1 = world
    pos: Tuple[float, float, float]
    joints: List[Joint] = field(default_factory=list)
    geoms: List[Geom] = field(default_factory=list)


@dataclass
class Actuator:
    joint: str
    gear: float


@dataclass
class RobotDescription:
    name: str
    bodies: List[Body]
    actuators: List[Actuator]
    timestep: float
    frame_skip: int
    ctrl_limit: float
    density: float = 1000.0
    total_mass: Optional[float] = None       # MuJoCo compiler settotalmass
    friction: float = 1.0
    gravity: float = 9.81
    # soft-contact / joint-limit parameters of this repo's engine (not MuJoCo's solver)
    contact_stiffness: float = 2.0e4         # N/m
    contact_damping: float = 1.5             # Hunt-Crossley coefficient c (s/m): damper = min(k*pen*c, d_max)
    contact_damping_max: float = 200.0       # N s/m cap of the normal damper (explicit-integration stability)
    friction_viscous: float = 150.0          # N s/m below the Coulomb cap
    limit_timeconst: float = 0.02            # spring-damper time constant of joint limits (MuJoCo solreflimit default)
    root_qpos0: Optional[Tuple[float, ...]] = None   # initial coordinates of the root joint(s) (default zeros / identity)
    contacts: bool = True                    # False: no geom collides (MuJoCo contype = conaffinity = 0, reacher.xml)


# ---------------------------------------------------------------------------------------------------------------
def half_cheetah() -> RobotDescription:
    """gym `half_cheetah.xml` (v3 env: frame_skip 5, timestep 0.01), planar: slide x, slide z, hinge y root."""
    dflt = dict(armature=0.1)

    def hinge(name, rng, stiff, damp):
        return Joint(name, "hinge", axis=(0, 1, 0), range=rng, stiffness=stiff, damping=damp, **dflt)

    r = 0.046
    bodies = [
        Body("torso", -1, (0, 0, 0.7),
             joints=[Joint("rootx", "slide", axis=(1, 0, 0)), Joint("rootz", "slide", axis=(0, 0, 1)),
                     Joint("rooty", "hinge", axis=(0, 1, 0))],
             geoms=[Geom("torso", "capsule", r, fromto=(-0.5, 0, 0, 0.5, 0, 0)),
                    Geom("head", "capsule", r, pos=(0.6, 0, 0.1), half=0.15, yangle=0.87)]),
        Body("bthigh", 0, (-0.5, 0, 0), joints=[hinge("bthigh", (-0.52, 1.05), 240, 6)],
             geoms=[Geom("bthigh", "capsule", r, pos=(0.1, 0, -0.13), half=0.145, yangle=-3.8)]),
        Body("bshin", 1, (0.16, 0, -0.25), joints=[hinge("bshin", (-0.785, 0.785), 180, 4.5)],
             geoms=[Geom("bshin", "capsule", r, pos=(-0.14, 0, -0.07), half=0.15, yangle=-2.03)]),
        Body("bfoot", 2, (-0.28, 0, -0.14), joints=[hinge("bfoot", (-0.4, 0.785), 120, 3)],
             geoms=[Geom("bfoot", "capsule", r, pos=(0.03, 0, -0.097), half=0.094, yangle=-0.27)]),
        Body("fthigh", 0, (0.5, 0, 0), joints=[hinge("fthigh", (-1.0, 0.7), 180, 4.5)],
             geoms=[Geom("fthigh", "capsule", r, pos=(-0.07, 0, -0.12), half=0.133, yangle=0.52)]),
        Body("fshin", 4, (-0.14, 0, -0.24), joints=[hinge("fshin", (-1.2, 0.87), 120, 3)],
             geoms=[Geom("fshin", "capsule", r, pos=(0.065, 0, -0.09), half=0.106, yangle=-0.6)]),
        Body("ffoot", 5, (0.13, 0, -0.18), joints=[hinge("ffoot", (-0.5, 0.5), 60, 1.5)],
             geoms=[Geom("ffoot", "capsule", r, pos=(0.045, 0, -0.07), half=0.07, yangle=-0.6)]),
    ]
    acts = [Actuator("bthigh", 120), Actuator("bshin", 90), Actuator("bfoot", 60), Actuator("fthigh", 120),
            Actuator("fshin", 60), Actuator("ffoot", 30)]
    return RobotDescription("halfcheetah", bodies, acts, timestep=0.01, frame_skip=5, ctrl_limit=1.0,
                            total_mass=14.0, friction=0.4, contact_stiffness=1.0e4, contact_damping=1.5,
                            contact_damping_max=60.0, friction_viscous=60.0)


def humanoid_standup() -> RobotDescription:
    """gym `humanoidstandup.xml` (timestep 0.003, frame_skip 5): free root + 17 hinges, 13 moving bodies; the pose
    at qpos0 is the standing humanoid rotated to lie on its back (see envs.humanoid_standup_qpos0)."""
    d2r = math.pi / 180.0

    def hinge(name, axis, pos, rng, stiff, damp, arm):
        return Joint(name, "hinge", axis=axis, pos=pos, range=(rng[0] * d2r, rng[1] * d2r), stiffness=stiff,
                     damping=damp, armature=arm)

    def leg(side, parent, s):          # s = -1 right, +1 left (sign of y)
        return [
            Body(f"{side}_thigh", parent, (0, 0.1 * s, -0.04),
                 joints=[hinge(f"{side}_hip_x", (-s * 1.0, 0, 0), (0, 0, 0), (-25, 5), 10, 5, 0.01),
                         hinge(f"{side}_hip_z", (0, 0, -s * 1.0), (0, 0, 0), (-60, 35), 10, 5, 0.01),
                         hinge(f"{side}_hip_y", (0, 1, 0), (0, 0, 0), (-110, 20), 20, 5, 0.008)],
                 geoms=[Geom(f"{side}_thigh1", "capsule", 0.06, fromto=(0, 0, 0, 0, -0.01 * s, -0.34))]),
            Body(f"{side}_shin", parent + 1, (0, -0.01 * s, -0.403),
                 joints=[hinge(f"{side}_knee", (0, -1, 0), (0, 0, 0.02), (-160, -2), 0, 1, 0.006)],
                 geoms=[Geom(f"{side}_shin1", "capsule", 0.049, fromto=(0, 0, 0, 0, 0, -0.3))]),
            Body(f"{side}_foot", parent + 2, (0, 0, -0.39),
                 geoms=[Geom(f"{side}_foot", "sphere", 0.075, pos=(0, 0, 0.1))]),
        ]

    def arm(side, parent, s):
        return [
            Body(f"{side}_upper_arm", 0, (0, 0.17 * s, 0.06),
                 joints=[hinge(f"{side}_shoulder1", (2, -s * 1.0, 1), (0, 0, 0), (-85, 60), 1, 1, 0.0068),
                         hinge(f"{side}_shoulder2", (0, s * 1.0, 1), (0, 0, 0), (-85, 60), 1, 1, 0.0051)],
                 geoms=[Geom(f"{side}_uarm1", "capsule", 0.04, fromto=(0, 0, 0, 0.16, 0.16 * s, -0.16))]),
            Body(f"{side}_lower_arm", parent, (0.18, 0.18 * s, -0.18),
                 joints=[hinge(f"{side}_elbow", (0, -1, -s * 1.0), (0, 0, 0), (-90, 50), 0, 1, 0.0028)],
                 geoms=[Geom(f"{side}_larm", "capsule", 0.031, fromto=(0.01, -0.01 * s, 0.01, 0.17, -0.17 * s, 0.17)),
                        Geom(f"{side}_hand", "sphere", 0.04, pos=(0.18, -0.18 * s, 0.18))]),
        ]

    bodies = [
        Body("torso", -1, (0, 0, 0), joints=[Joint("root", "free")],
             geoms=[Geom("torso1", "capsule", 0.07, fromto=(0, -0.07, 0, 0, 0.07, 0)),
                    Geom("head", "sphere", 0.09, pos=(0, 0, 0.19)),
                    Geom("uwaist", "capsule", 0.06, fromto=(-0.01, -0.06, -0.12, -0.01, 0.06, -0.12))]),
        Body("lwaist", 0, (-0.01, 0, -0.26),
             joints=[hinge("abdomen_z", (0, 0, 1), (0, 0, 0.065), (-45, 45), 20, 5, 0.02),
                     hinge("abdomen_y", (0, 1, 0), (0, 0, 0.065), (-75, 30), 10, 5, 0.02)],
             geoms=[Geom("lwaist", "capsule", 0.06, fromto=(0, -0.06, 0, 0, 0.06, 0))]),
        Body("pelvis", 1, (0, 0, -0.165),
             joints=[hinge("abdomen_x", (1, 0, 0), (0, 0, 0.1), (-35, 35), 10, 5, 0.02)],
             geoms=[Geom("butt", "capsule", 0.09, fromto=(-0.02, -0.07, 0, -0.02, 0.07, 0))]),
    ]
    bodies += leg("right", 2, -1.0)      # bodies 3,4,5
    bodies += leg("left", 2, +1.0)       # bodies 6,7,8  (parents fixed below)
    bodies += arm("right", 9, -1.0)      # bodies 9,10
    bodies += arm("left", 11, +1.0)      # bodies 11,12
    # fix the chain parents (leg() numbers shin/foot relative to `parent`)
    bodies[3].parent, bodies[4].parent, bodies[5].parent = 2, 3, 4
    bodies[6].parent, bodies[7].parent, bodies[8].parent = 2, 6, 7
    bodies[9].parent, bodies[10].parent = 0, 9
    bodies[11].parent, bodies[12].parent = 0, 11
    acts = [Actuator("abdomen_y", 100), Actuator("abdomen_z", 100), Actuator("abdomen_x", 100),
            Actuator("right_hip_x", 100), Actuator("right_hip_z", 100), Actuator("right_hip_y", 300),
            Actuator("right_knee", 200), Actuator("left_hip_x", 100), Actuator("left_hip_z", 100),
            Actuator("left_hip_y", 300), Actuator("left_knee", 200), Actuator("right_shoulder1", 25),
            Actuator("right_shoulder2", 25), Actuator("right_elbow", 25), Actuator("left_shoulder1", 25),
            Actuator("left_shoulder2", 25), Actuator("left_elbow", 25)]
    return RobotDescription("humanoid_standup", bodies, acts, timestep=0.003, frame_skip=5, ctrl_limit=0.4,
                            friction=1.0, contact_stiffness=2.0e4, contact_damping=1.5, friction_viscous=150.0,
                            # lying on the back: root at z = 0.105, rotated -90 deg about y (head towards -x, face up)
                            root_qpos0=(0.0, 0.0, 0.105, math.sqrt(0.5), 0.0, -math.sqrt(0.5), 0.0))


def hopper() -> RobotDescription:
    """gym `hopper.xml` (Hopper-v3: timestep 0.002, frame_skip 4), planar: slide x, slide z, hinge y root, then
    thigh / leg / foot hinges about -y.  The XML is written in global coordinates (torso at z = 1.25); here every body
    frame sits at its joint and the root slide carries the height, so qpos[1] is the absolute torso height the
    reference's `unhealthy_states` reads (environments/mujoco.py:196-212)."""
    d2r = math.pi / 180.0

    def hinge(name, pos, rng):
        return Joint(name, "hinge", axis=(0, -1, 0), pos=pos, range=(rng[0] * d2r, rng[1] * d2r), damping=1.0,
                     armature=1.0)

    bodies = [
        Body("torso", -1, (0, 0, 0),
             joints=[Joint("rootx", "slide", axis=(1, 0, 0)), Joint("rootz", "slide", axis=(0, 0, 1)),
                     Joint("rooty", "hinge", axis=(0, 1, 0))],
             geoms=[Geom("torso", "capsule", 0.05, fromto=(0, 0, 0.2, 0, 0, -0.2))]),
        Body("thigh", 0, (0, 0, -0.2), joints=[hinge("thigh_joint", (0, 0, 0), (-150, 0))],
             geoms=[Geom("thigh", "capsule", 0.05, fromto=(0, 0, 0, 0, 0, -0.45))]),
        Body("leg", 1, (0, 0, -0.7), joints=[hinge("leg_joint", (0, 0, 0.25), (-150, 0))],
             geoms=[Geom("leg", "capsule", 0.04, fromto=(0, 0, 0.25, 0, 0, -0.25))]),
        Body("foot", 2, (0.065, 0, -0.25), joints=[hinge("foot_joint", (-0.065, 0, 0), (-45, 45))],
             geoms=[Geom("foot", "capsule", 0.06, fromto=(-0.195, 0, 0, 0.195, 0, 0))]),
    ]
    acts = [Actuator("thigh_joint", 200), Actuator("leg_joint", 200), Actuator("foot_joint", 200)]
    return RobotDescription("hopper", bodies, acts, timestep=0.002, frame_skip=4, ctrl_limit=1.0, friction=0.9,
                            contact_stiffness=2.0e4, contact_damping=1.5, contact_damping_max=120.0,
                            friction_viscous=120.0, root_qpos0=(0.0, 1.25, 0.0))


def ant() -> RobotDescription:
    """gym `ant.xml` (Ant-v3: timestep 0.01, frame_skip 5): free root (sphere torso) and four two-joint legs (hip about
    z, ankle about a horizontal axis); density 5, armature 1, damping 1, gear 150.  The XML's jointless leg-root
    bodies are merged into the torso (their capsules become torso geoms)."""
    d2r = math.pi / 180.0

    def hinge(name, axis, rng):
        return Joint(name, "hinge", axis=axis, range=(rng[0] * d2r, rng[1] * d2r), damping=1.0, armature=1.0)

    legs = [  # name, (sx, sy), ankle axis, ankle range
        ("1", (1, 1), (-1, 1, 0), (30, 70)), ("2", (-1, 1), (1, 1, 0), (-70, -30)),
        ("3", (-1, -1), (-1, 1, 0), (-70, -30)), ("4", (1, -1), (1, 1, 0), (30, 70))]
    torso_geoms = [Geom("torso", "sphere", 0.25)]
    bodies = [Body("torso", -1, (0, 0, 0), joints=[Joint("root", "free")], geoms=torso_geoms)]
    for n, (sx, sy), axis, rng in legs:
        torso_geoms.append(Geom(f"aux_{n}", "capsule", 0.08, fromto=(0, 0, 0, 0.2 * sx, 0.2 * sy, 0)))
        hip = len(bodies)
        bodies.append(Body(f"aux_{n}", 0, (0.2 * sx, 0.2 * sy, 0), joints=[hinge(f"hip_{n}", (0, 0, 1), (-30, 30))],
                           geoms=[Geom(f"leg_{n}", "capsule", 0.08, fromto=(0, 0, 0, 0.2 * sx, 0.2 * sy, 0))]))
        bodies.append(Body(f"ankle_{n}", hip, (0.2 * sx, 0.2 * sy, 0), joints=[hinge(f"ankle_{n}", axis, rng)],
                           geoms=[Geom(f"ankle_{n}", "capsule", 0.08, fromto=(0, 0, 0, 0.4 * sx, 0.4 * sy, 0))]))
    acts = [Actuator(j, 150) for j in ("hip_4", "ankle_4", "hip_1", "ankle_1", "hip_2", "ankle_2", "hip_3", "ankle_3")]
    return RobotDescription("ant", bodies, acts, timestep=0.01, frame_skip=5, ctrl_limit=1.0, density=5.0,
                            friction=1.0, contact_stiffness=1.0e3, contact_damping=1.5, contact_damping_max=20.0,
                            friction_viscous=20.0, root_qpos0=(0.0, 0.0, 0.75, 1.0, 0.0, 0.0, 0.0))


def humanoid() -> RobotDescription:
    """gym `humanoid.xml` (Humanoid-v3): the same body, joints and gears as `humanoidstandup.xml`, starting upright with
    the root at z = 1.4 instead of lying on its back."""
    d = humanoid_standup()
    d.name = "humanoid"
    d.root_qpos0 = (0.0, 0.0, 1.4, 1.0, 0.0, 0.0, 0.0)
    return d


ROBOTS = {"halfcheetah": half_cheetah, "humanoid_standup": humanoid_standup, "hopper": hopper, "ant": ant,
          "humanoid": humanoid}


def reacher() -> RobotDescription:
    """gym `reacher.xml` (Reacher-v2: timestep 0.01, frame_skip 2): a planar two-link arm on hinges about z (default
    joint armature 1, damping 1; joint0 unlimited, joint1 in [-3, 3] rad; motors gear 200, ctrl in [-1, 1]), the
    fingertip 0.11 beyond the second joint, and the target as its own body on two undriven slide joints (x, y) -- it
    keeps whatever position the start state gives it.  No geom collides (contype = conaffinity = 0)."""
    j = dict(armature=1.0, damping=1.0)
    bodies = [
        Body("body0", -1, (0, 0, 0.01), joints=[Joint("joint0", "hinge", axis=(0, 0, 1), **j)],
             geoms=[Geom("link0", "capsule", 0.01, fromto=(0, 0, 0, 0.1, 0, 0))]),
        Body("body1", 0, (0.1, 0, 0), joints=[Joint("joint1", "hinge", axis=(0, 0, 1), range=(-3.0, 3.0), **j)],
             geoms=[Geom("link1", "capsule", 0.01, fromto=(0, 0, 0, 0.1, 0, 0)),
                    Geom("fingertip", "sphere", 0.01, pos=(0.11, 0, 0))]),
        # reacher.xml places the body at (.1, -.1, .01) and gives the slides ref = .1 / -.1: the joint coordinates ARE the
        # target's world x, y.  The slides carry no force (no damping, no actuator, gravity is along z); armature 1
        # instead of the XML's 0 keeps the 3-milligram body out of the fp32 factorisation -- its acceleration is 0 either way.
        Body("target", -1, (0.0, 0.0, 0.01),
             joints=[Joint("target_x", "slide", axis=(1, 0, 0), range=(-0.27, 0.27), armature=1.0),
                     Joint("target_y", "slide", axis=(0, 1, 0), range=(-0.27, 0.27), armature=1.0)],
             geoms=[Geom("target", "sphere", 0.009)]),
    ]
    return RobotDescription("reacher", bodies, [Actuator("joint0", 200.0), Actuator("joint1", 200.0)], timestep=0.01,
                            frame_skip=2, ctrl_limit=1.0, contacts=False)


# ---------------------------------------------------------------------------------------------------------------
ROBOTS["reacher"] = reacher


def _capsule_mass_inertia(radius, length, density):
    """Solid capsule along z about its centre: (mass, Ixx(=Iyy), Izz)."""
    r, L = radius, length
    m_cyl = density * math.pi * r * r * L
    m_hemi = density * (2.0 / 3.0) * math.pi * r ** 3
    izz = 0.5 * m_cyl * r * r + 2 * m_hemi * 0.4 * r * r
    ixx = m_cyl * (L * L / 12.0 + r * r / 4.0) + 2 * m_hemi * (0.4 * r * r + 0.25 * L * L + 0.375 * r * L)
    return m_cyl + 2 * m_hemi, ixx, izz


def _geom_inertia(g: Geom, density):
    """(mass, centre[3], inertia 3x3 about the centre, body-frame axes)."""
    if g.kind == "sphere":
        m = density * (4.0 / 3.0) * math.pi * g.size ** 3
        return m, np.asarray(g.pos, float), 0.4 * m * g.size ** 2 * np.eye(3)
    a, b = g.endpoints()
    L = float(np.linalg.norm(b - a))
    z = (b - a) / L
    m, ixx, izz = _capsule_mass_inertia(g.size, L, density)
    inertia = ixx * (np.eye(3) - np.outer(z, z)) + izz * np.outer(z, z)
    return m, 0.5 * (a + b), inertia


@dataclass
class CompiledModel:
    """Flat tables (float64 here; the C ABI takes float32 copies).  Index conventions in include/icem_b200.h."""
    name: str
    nb: int
    nq: int
    nv: int
    nu: int
    nc: int
    dt: float
    nsub: int
    gravity: float
    ctrl_limit: float
    contact_stiffness: float
    contact_damping: float
    contact_damping_max: float
    friction_viscous: float
    friction: float
    # bodies
    body_parent: np.ndarray       # [nb] int
    body_depth: np.ndarray        # [nb] int
    body_pos: np.ndarray          # [nb,3]
    body_mass: np.ndarray         # [nb]
    body_com: np.ndarray          # [nb,3]
    body_inertia: np.ndarray      # [nb,6] xx yy zz xy xz yz about the com, body axes
    body_dof_start: np.ndarray    # [nb] int
    body_dof_count: np.ndarray    # [nb] int
    # dofs
    dof_body: np.ndarray          # [nv] int
    dof_type: np.ndarray          # [nv] int (SLIDE / HINGE / FREE_TRANS / FREE_ROT)
    dof_axis: np.ndarray          # [nv,3] unit, body frame
    dof_anchor: np.ndarray        # [nv,3] body frame
    dof_qadr: np.ndarray          # [nv] int: index of the coordinate in qpos (FREE_ROT: start of the quaternion)
    dof_parent: np.ndarray        # [nv] int: previous dof on the path to the root, -1 for none
    dof_stiffness: np.ndarray
    dof_damping: np.ndarray
    dof_armature: np.ndarray
    dof_limited: np.ndarray       # [nv] int
    dof_lo: np.ndarray
    dof_hi: np.ndarray
    dof_klim: np.ndarray          # limit spring (N m / rad)
    dof_blim: np.ndarray          # limit damper
    dof_gear: np.ndarray          # [nv] gear of the actuator on this dof (0 = unactuated)
    dof_act: np.ndarray           # [nv] int actuator (control) index, -1 = unactuated
    # contact points (spheres fixed to bodies against the plane z = 0)
    con_body: np.ndarray          # [nc] int
    con_pos: np.ndarray           # [nc,3]
    con_radius: np.ndarray        # [nc]
    qpos0: np.ndarray             # [nq]
    integrator: str = "euler"     # "euler" (semi-implicit, springs / dampers implicit) | "rk4" (mj_RungeKutta, N = 4)

    @property
    def max_depth(self):
        return int(self.body_depth.max())


def compile_model(desc: RobotDescription) -> CompiledModel:
    nb = len(desc.bodies)
    assert nb <= MAX_BODIES
    parent = np.array([b.parent for b in desc.bodies], np.int64)
    assert all(parent[i] < i for i in range(nb)), "bodies must be listed parent-before-child"
    depth = np.zeros(nb, np.int64)
    for i in range(nb):
        depth[i] = 0 if parent[i] < 0 else depth[parent[i]] + 1
    nchild = np.zeros(nb, np.int64)
    for i in range(nb):
        if parent[i] >= 0:
            nchild[parent[i]] += 1
    assert nchild.max() <= MAX_CHILDREN

    # mass properties from the geoms
    mass = np.zeros(nb)
    com = np.zeros((nb, 3))
    inertia = np.zeros((nb, 3, 3))
    for i, b in enumerate(desc.bodies):
        parts = [_geom_inertia(g, desc.density) for g in b.geoms]
        m = sum(p[0] for p in parts)
        c = sum(p[0] * p[1] for p in parts) / m
        I = np.zeros((3, 3))
        for pm, pc, pI in parts:
            r = pc - c
            I += pI + pm * (np.dot(r, r) * np.eye(3) - np.outer(r, r))
        mass[i], com[i], inertia[i] = m, c, I
    if desc.total_mass is not None:
        s = desc.total_mass / mass.sum()
        mass *= s
        inertia *= s

    # dofs
    rows = []
    qadr = 0
    dof_start = np.zeros(nb, np.int64)
    dof_count = np.zeros(nb, np.int64)
    last_dof_of_body = -np.ones(nb, np.int64)
    qpos0 = []
    joint_first_dof = {}
    for i, b in enumerate(desc.bodies):
        dof_start[i] = len(rows)
        prev = last_dof_of_body[parent[i]] if parent[i] >= 0 else -1
        # bodies without joints inherit the chain end of their parent
        for j in b.joints:
            joint_first_dof[j.name] = len(rows)
            if j.kind == "free":
                for a in range(3):
                    rows.append(dict(body=i, type=FREE_TRANS, axis=np.eye(3)[a], anchor=np.zeros(3), qadr=qadr + a,
                                     parent=prev, joint=j))
                    prev = len(rows) - 1
                for a in range(3):
                    rows.append(dict(body=i, type=FREE_ROT, axis=np.eye(3)[a], anchor=np.zeros(3), qadr=qadr + 3,
                                     parent=prev, joint=j))
                    prev = len(rows) - 1
                qpos0 += [0, 0, 0, 1, 0, 0, 0]
                qadr += 7
            else:
                ax = np.asarray(j.axis, float)
                ax = ax / np.linalg.norm(ax)
                rows.append(dict(body=i, type=SLIDE if j.kind == "slide" else HINGE, axis=ax,
                                 anchor=np.asarray(j.pos, float), qadr=qadr, parent=prev, joint=j))
                prev = len(rows) - 1
                qpos0.append(0.0)
                qadr += 1
        dof_count[i] = len(rows) - dof_start[i]
        last_dof_of_body[i] = prev
    nv, nq = len(rows), qadr
    assert nv <= MAX_DOFS

    gear = np.zeros(nv)
    act = -np.ones(nv, np.int64)
    for u, a in enumerate(desc.actuators):
        d = joint_first_dof[a.joint]
        gear[d], act[d] = a.gear, u

    # contact points: capsule end spheres and spheres
    cb, cp, cr = [], [], []
    for i, b in enumerate(desc.bodies):
        for g in b.geoms:
            for e in g.endpoints():
                if desc.contacts:
                    cb.append(i); cp.append(e); cr.append(g.size)
    nc = len(cb)
    assert nc <= MAX_CONTACTS, nc

    m = CompiledModel(
        name=desc.name, nb=nb, nq=nq, nv=nv, nu=len(desc.actuators), nc=nc, dt=desc.timestep, nsub=desc.frame_skip,
        gravity=desc.gravity, ctrl_limit=desc.ctrl_limit, contact_stiffness=desc.contact_stiffness,
        contact_damping=desc.contact_damping, contact_damping_max=desc.contact_damping_max, friction_viscous=desc.friction_viscous, friction=desc.friction,
        body_parent=parent, body_depth=depth, body_pos=np.array([b.pos for b in desc.bodies], float),
        body_mass=mass, body_com=com,
        body_inertia=np.stack([inertia[:, 0, 0], inertia[:, 1, 1], inertia[:, 2, 2], inertia[:, 0, 1],
                               inertia[:, 0, 2], inertia[:, 1, 2]], axis=1),
        body_dof_start=dof_start, body_dof_count=dof_count,
        dof_body=np.array([r["body"] for r in rows], np.int64), dof_type=np.array([r["type"] for r in rows], np.int64),
        dof_axis=np.array([r["axis"] for r in rows], float), dof_anchor=np.array([r["anchor"] for r in rows], float),
        dof_qadr=np.array([r["qadr"] for r in rows], np.int64), dof_parent=np.array([r["parent"] for r in rows], np.int64),
        dof_stiffness=np.array([r["joint"].stiffness if r["type"] in (SLIDE, HINGE) else 0.0 for r in rows]),
        dof_damping=np.array([r["joint"].damping if r["type"] in (SLIDE, HINGE) else 0.0 for r in rows]),
        dof_armature=np.array([r["joint"].armature if r["type"] in (SLIDE, HINGE) else 0.0 for r in rows]),
        dof_limited=np.array([int(r["joint"].range is not None) for r in rows], np.int64),
        dof_lo=np.array([r["joint"].range[0] if r["joint"].range is not None else 0.0 for r in rows]),
        dof_hi=np.array([r["joint"].range[1] if r["joint"].range is not None else 0.0 for r in rows]),
        dof_klim=np.zeros(nv), dof_blim=np.zeros(nv), dof_gear=gear, dof_act=act,
        con_body=np.array(cb, np.int64), con_pos=np.array(cp, float).reshape(-1, 3), con_radius=np.array(cr, float),
        qpos0=np.array(qpos0, float))
    if desc.root_qpos0 is not None:
        m.qpos0[:len(desc.root_qpos0)] = desc.root_qpos0
    _limit_gains(m, desc.limit_timeconst)
    return m


def _limit_gains(m: CompiledModel, timeconst):
    """Joint-limit spring/damper from the reference-pose joint-space inertia (critically damped, time constant
    `timeconst`): k = M_jj / tc^2, b = 2 M_jj / tc.  M_jj at qpos0 for a hinge = armature + sum over the bodies it
    moves of (axis . I_b axis + m_b * dist(com_b, axis)^2); joints at zero => world frame = sum of body offsets."""
    origin = np.zeros((m.nb, 3))
    for i in range(m.nb):
        origin[i] = m.body_pos[i] + (origin[m.body_parent[i]] if m.body_parent[i] >= 0 else 0.0)
    moved = [[] for _ in range(m.nv)]          # bodies in the subtree of each dof's body
    for b in range(m.nb):
        a = b
        chain = set()
        while a >= 0:
            chain.add(a)
            a = m.body_parent[a]
        for j in range(m.nv):
            if m.dof_body[j] in chain:
                moved[j].append(b)
    for j in range(m.nv):
        if not m.dof_limited[j]:
            continue
        ax = m.dof_axis[j]
        anchor = origin[m.dof_body[j]] + m.dof_anchor[j]
        mjj = m.dof_armature[j]
        for b in moved[j]:
            I = m.body_inertia[b]
            Im = np.array([[I[0], I[3], I[4]], [I[3], I[1], I[5]], [I[4], I[5], I[2]]])
            r = origin[b] + m.body_com[b] - anchor
            if m.dof_type[j] == HINGE:
                perp = r - np.dot(r, ax) * ax
                mjj += ax @ Im @ ax + m.body_mass[b] * np.dot(perp, perp)
            else:
                mjj += m.body_mass[b]
        m.dof_klim[j] = mjj / timeconst ** 2
        m.dof_blim[j] = 2.0 * mjj / timeconst


_CACHE = {}


def get_model(name, integrator="euler") -> CompiledModel:
    """`integrator="rk4"`: the Runge-Kutta substep gym's XML files ask of MuJoCo."""
    key = (name, integrator)
    if key not in _CACHE:
        import dataclasses
        m = compile_model(ROBOTS[name]())
        _CACHE[key] = dataclasses.replace(m, integrator=integrator)
    return _CACHE[key]
