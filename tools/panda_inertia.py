"""Reflected inertia of the seven Panda arm joints at the initial pose, from the reference's own assets.

franka_panda.urdf has no <inertial> blocks, so IsaacGym derives every link's mass properties from its collision mesh
at the default density (1000 kg/m^3; isaacgym_wrapper.py never sets one). This script does the same from
meshes/collision/*.obj (volume, centre of mass and inertia tensor of the closed triangle mesh by signed tetrahedra),
places the links with the URDF chain (franka_panda.urdf:27-242) at q0 (panda_env/panda.yaml:10) and prints, for each
joint, the inertia of everything distal to it about the joint axis: the `joint_inertia` the velocity drive of the
integrator works against (m3p2i_b200/scene.py). Needs /root/reference; run in the build container:
    python tools/panda_inertia.py
"""
import os
import numpy as np

MESH = "/root/reference/src/m3p2i_aip/assets/urdf/franka_description/meshes/collision"
DENSITY = 1000.0
XYZ = [(0, 0, 0.333), (0, 0, 0), (0, -0.316, 0), (0.0825, 0, 0), (-0.0825, 0.384, 0), (0, 0, 0), (0.088, 0, 0)]
ROLL = [0, -1, 1, 1, -1, 1, 1]  # multiples of pi/2
Q0 = [0, 0, 0, -2, 0, 1.8675, 0]


def mesh_props(name):
    v, f = [], []
    for line in open(os.path.join(MESH, name)):
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            v.append([float(x) for x in p[1:4]])
        elif p[0] == "f":
            idx = [int(t.split("/")[0]) - 1 for t in p[1:]]
            for i in range(1, len(idx) - 1):
                f.append([idx[0], idx[i], idx[i + 1]])
    v, f = np.array(v), np.array(f)
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    vol6 = np.einsum("ij,ij->i", a, np.cross(b, c))
    vol = vol6.sum() / 6.0
    com = ((a + b + c) * vol6[:, None]).sum(0) / (24.0 * vol)
    # second moments about the origin: integral of x_i x_j over each tetrahedron (0, a, b, c)
    S = np.zeros((3, 3))
    for p_, q_ in ((a, a), (b, b), (c, c)):
        S += np.einsum("i,ij,ik->jk", vol6, p_, q_) * 2
    for p_, q_ in ((a, b), (a, c), (b, c)):
        S += np.einsum("i,ij,ik->jk", vol6, p_, q_) + np.einsum("i,ij,ik->jk", vol6, q_, p_)
    S /= 120.0
    if vol < 0:
        vol, S = -vol, -S
    m = DENSITY * vol
    S = DENSITY * S
    I0 = np.trace(S) * np.eye(3) - S                      # inertia about the mesh origin
    Ic = I0 - m * (com @ com * np.eye(3) - np.outer(com, com))
    return m, com, Ic


def rx(k):
    c, s = [(1, 0), (0, 1), (-1, 0), (0, -1)][k % 4]
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], float)


def rz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def main():
    R, p = np.eye(3), np.zeros(3)
    bodies, axes = [], []     # (first joint that moves it, mass, world com, world inertia about com)
    for j in range(7):
        p = p + R @ np.array(XYZ[j])
        R = R @ rx(ROLL[j])
        axes.append((p.copy(), R[:, 2].copy()))
        R = R @ rz(Q0[j])
        m, com, Ic = mesh_props(f"link{j + 1}.obj")
        bodies.append((j, m, p + R @ com, R @ Ic @ R.T))
    ph = p + R @ np.array([0, 0, 0.107])
    Rh = R @ rz(-np.pi / 4)
    m, com, Ic = mesh_props("hand.obj")
    bodies.append((6, m, ph + Rh @ com, Rh @ Ic @ Rh.T))
    m, com, Ic = mesh_props("finger.obj")
    for sg in (1.0, -1.0):
        Rf = Rh @ np.diag([sg, sg, 1.0])
        bodies.append((6, m, ph + Rh @ np.array([0, sg * 0.02, 0.0584]) + Rf @ com, Rf @ Ic @ Rf.T))
    print("link masses (kg):", [round(b[1], 3) for b in bodies], "total", round(sum(b[1] for b in bodies), 2))
    out = []
    for j, (o, ax) in enumerate(axes):
        I = 0.0
        for first, m, c, Ic in bodies:
            if first < j:
                continue
            d = c - o
            perp = d - (d @ ax) * ax
            I += ax @ Ic @ ax + m * (perp @ perp)
        out.append(I)
    print("reflected inertia about each joint axis at q0 (kg m^2):", [round(x, 4) for x in out])


if __name__ == "__main__":
    main()
