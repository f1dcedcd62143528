"""ctypes wrapper of tests/emu/libteam_emu.so: the lane-cooperative Panda rollout kernel's device code compiled for
the host and run in lock step (cuda_emu.h). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from m3p2i_b200 import _abi as A

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "m3p2i-aip_b200", "csrc")
SO = os.path.join(HERE, "libteam_emu.so")
_LIB = None


def build(force=False):
    deps = [os.path.join(HERE, f) for f in ("team_emu.cpp", "cuda_emu.h")] + \
           [os.path.join(CSRC, f) for f in ("common.cuh", "panda_env.cuh", "rollout_common.cuh", "panda_team.cuh", "panda_far.cuh", "params_host.h")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unused", "-o", SO,
                               os.path.join(HERE, "team_emu.cpp"), "-I", os.path.join(ROOT, "include")])
    return SO


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.emu_team_rollout_actions.restype = C.c_int
        L.emu_team_rollout_actions.argtypes = [C.POINTER(A.Config), C.POINTER(A.PandaScene), C.c_int, A.fp, C.c_int, A.fp, A.fp, A.fp,
                                               C.c_int, C.c_int, A.fp, A.fp, A.fp, C.POINTER(C.c_long)]
        L.emu_thread_rollout_actions.restype = C.c_int
        L.emu_thread_rollout_actions.argtypes = [C.POINTER(A.Config), C.POINTER(A.PandaScene), C.c_int, A.fp, C.c_int, A.fp, A.fp,
                                                 A.fp, A.fp, A.fp, A.fp]
        L.emu_far_rollout_actions.restype = C.c_int
        L.emu_far_rollout_actions.argtypes = [C.POINTER(A.Config), C.POINTER(A.PandaScene), C.c_int, A.fp, C.c_int, A.fp, A.fp, A.fp,
                                              C.POINTER(C.c_int), C.POINTER(C.c_int), A.fp, A.fp, A.fp, A.fp]
        L.emu_split_rollout_actions.restype = C.c_int
        L.emu_split_rollout_actions.argtypes = [C.POINTER(A.Config), C.POINTER(A.PandaScene), C.c_int, A.fp, C.c_int, A.fp, A.fp, A.fp,
                                                C.c_int, C.c_int, A.fp, A.fp, A.fp, A.fp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


def rollout_actions(config, scene, task, goal, gripper, dof, root, actions, lanes, block_threads=64):
    """-> states [K,T,4], cost_horizon [K,T], env_end [K,53], number of warp collectives executed"""
    K, T = config.num_samples, config.horizon
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    actions, dof, root, goal = f(actions), f(dof).ravel(), f(root).ravel(), f(goal).ravel()
    assert actions.shape == (K, T, 9)
    g7 = np.zeros(8, np.float32)
    g7[:goal.size] = goal
    st, ch, env = np.zeros((K, T, 4), np.float32), np.zeros((K, T), np.float32), np.zeros((K, 53), np.float32)
    n = C.c_long(0)
    rc = lib().emu_team_rollout_actions(C.byref(config), C.byref(scene), A.TASK_IDS[task], A.as_fp(g7), A.GRIPPER_IDS[gripper],
                                        A.as_fp(dof), A.as_fp(root), A.as_fp(actions), lanes, block_threads, A.as_fp(st),
                                        A.as_fp(ch), A.as_fp(env), C.byref(n))
    if rc:
        raise RuntimeError(f"emu_team_rollout_actions rc={rc}")
    return st, ch, env, n.value


def thread_rollout_actions(config, scene, task, goal, gripper, dof, root, actions):
    """The same rollout through the thread-per-sample device code (pick / place) -> states, cost_horizon, env_end"""
    K, T = config.num_samples, config.horizon
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    actions, dof, root, goal = f(actions), f(dof).ravel(), f(root).ravel(), f(goal).ravel()
    g7 = np.zeros(8, np.float32)
    g7[:goal.size] = goal
    st, ch, env = np.zeros((K, T, 4), np.float32), np.zeros((K, T), np.float32), np.zeros((K, 53), np.float32)
    rc = lib().emu_thread_rollout_actions(C.byref(config), C.byref(scene), A.TASK_IDS[task], A.as_fp(g7), A.GRIPPER_IDS[gripper],
                                          A.as_fp(dof), A.as_fp(root), A.as_fp(actions), A.as_fp(st), A.as_fp(ch), A.as_fp(env))
    if rc:
        raise RuntimeError(f"emu_thread_rollout_actions rc={rc}")
    return st, ch, env


def far_rollout_actions(config, scene, task, goal, gripper, dof, root, actions):
    """The far-field kernel's device code (panda_far.cuh) on the same open-loop rollout -> ok [K] (sample stayed in the far
    field), prod_ok [2] (rows 0 / K/2 of the batch), states [K,T,4], cost_horizon [K,T], cost_sum [K], J [K]"""
    K, T = config.num_samples, config.horizon
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    actions, dof, root, goal = f(actions), f(dof).ravel(), f(root).ravel(), f(goal).ravel()
    g7 = np.zeros(8, np.float32)
    g7[:goal.size] = goal
    ok, pok = np.zeros(K, np.int32), np.zeros(2, np.int32)
    st, ch = np.zeros((K, T, 4), np.float32), np.zeros((K, T), np.float32)
    cs, J = np.zeros(K, np.float32), np.zeros(K, np.float32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rc = lib().emu_far_rollout_actions(C.byref(config), C.byref(scene), A.TASK_IDS[task], A.as_fp(g7), A.GRIPPER_IDS[gripper],
                                       A.as_fp(dof), A.as_fp(root), A.as_fp(actions), ip(ok), ip(pok), A.as_fp(st), A.as_fp(ch),
                                       A.as_fp(cs), A.as_fp(J))
    if rc:
        raise RuntimeError(f"emu_far_rollout_actions rc={rc}")
    return ok.astype(bool), pok.astype(bool), st, ch, cs, J


def split_rollout_actions(config, scene, task, goal, gripper, dof, root, actions, lanes, block_threads=64):
    """What a pick / place command launches: the far-field code over all samples, then the team kernel over the near list
    with its hand-over boundaries -> states [K,T,4], cost_horizon [K,T], cost_sum [K], J [K], far [K] (bool), boundary [K]"""
    K, T = config.num_samples, config.horizon
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    actions, dof, root, goal = f(actions), f(dof).ravel(), f(root).ravel(), f(goal).ravel()
    g7 = np.zeros(8, np.float32)
    g7[:goal.size] = goal
    far, bd = np.zeros(K, np.int32), np.zeros(K, np.int32)
    st, ch = np.zeros((K, T, 4), np.float32), np.zeros((K, T), np.float32)
    cs, J = np.zeros(K, np.float32), np.zeros(K, np.float32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rc = lib().emu_split_rollout_actions(C.byref(config), C.byref(scene), A.TASK_IDS[task], A.as_fp(g7), A.GRIPPER_IDS[gripper],
                                         A.as_fp(dof), A.as_fp(root), A.as_fp(actions), lanes, block_threads, A.as_fp(st),
                                         A.as_fp(ch), A.as_fp(cs), A.as_fp(J), ip(far), ip(bd))
    if rc:
        raise RuntimeError(f"emu_split_rollout_actions rc={rc}")
    return st, ch, cs, J, far.astype(bool), bd
