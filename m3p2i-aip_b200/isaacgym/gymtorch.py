"""isaacgym.gymtorch stand-in: tensors are plain torch tensors here."""


def wrap_tensor(t):
    return t


def unwrap_tensor(t):
    return t
