"""ctypes access to tests/emu (the kernel SOURCE compiled for the host).  Test-side only."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
# S2D_EMU_LIB: another build of the same source, e.g. one made with -fsanitize=undefined (see tests/emu/Makefile `ubsan`)
SO = os.environ.get("S2D_EMU_LIB") or os.path.join(EMU_DIR, "_build", "libemu_reachball.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = glob.glob(os.path.join(HERE, "..", "gym-soccer-2d-env_b200", "csrc", "*.cuh")) + \
            glob.glob(os.path.join(EMU_DIR, "*.cpp")) + glob.glob(os.path.join(EMU_DIR, "*.h"))
        if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(s) for s in srcs):
            subprocess.run(["make", "-s", "-B", "-C", EMU_DIR], check=True, capture_output=True)
        L = C.CDLL(SO)
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_void_p]
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_state.restype = C.c_void_p
        L.emu_state.argtypes = [C.c_void_p]
        L.emu_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.emu_uses_default_sp.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class EmuSim:
    """Same interface as oracle_lib.OracleSim, but the arithmetic is csrc/*.cuh compiled by g++."""

    def __init__(self, cfg):
        self.L = lib()
        self.n = int(cfg.num_envs)
        self.h = self.L.emu_create(C.byref(cfg))
        self.obs = np.zeros((self.n, 10), np.float32)
        self.term_obs = np.zeros((self.n, 10), np.float32)
        self.reward = np.zeros(self.n, np.float32)
        self.done = np.zeros(self.n, np.uint8)
        self.result = np.zeros(self.n, np.uint8)
        self.stats6 = np.zeros(6, np.float64)

    def close(self):
        if self.h is not None:
            self.L.emu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.emu_reset(self.h, _ptr(m), _ptr(self.obs))
        return self.obs

    def step(self, actions, k=1):
        a = np.ascontiguousarray(actions)
        self.L.emu_step(self.h, _ptr(a), int(k), _ptr(self.obs), _ptr(self.reward), _ptr(self.done), _ptr(self.result),
                        _ptr(self.term_obs), _ptr(self.stats6))
        return self.obs, self.reward, self.done, self.result

    def state_bytes(self):
        buf = (C.c_ubyte * (self.n * 80)).from_address(self.L.emu_state(self.h))
        return np.frombuffer(buf, dtype=np.uint8)

    def get_state(self):
        """[N, 20] float64 in oracle_lib.STATE_FIELDS order, from the plane-major layout."""
        raw = self.state_bytes()
        n = self.n
        f = raw[: 4 * n * 16].view(np.float32).reshape(4, n, 4).astype(np.float64)
        u = raw[4 * n * 16:].view(np.int32).reshape(n, 4).astype(np.float64)
        out = np.zeros((n, 20))
        out[:, 0:4] = f[0]
        out[:, 4:8] = f[1]
        out[:, 8] = f[3][:, 2]
        out[:, 9:13] = f[2]
        out[:, 13:15] = f[3][:, 0:2]
        out[:, 15] = f[3][:, 3]
        out[:, 16:20] = u
        return out

    def set_state(self, i, v):
        raw = self.state_bytes()
        n = self.n
        f = raw[: 4 * n * 16].view(np.float32).reshape(4, n, 4)
        u = raw[4 * n * 16:].view(np.int32).reshape(n, 4)
        v = [float(x) for x in v]
        f[0, i] = v[0:4]
        f[1, i] = v[4:8]
        f[2, i] = v[9:13]
        f[3, i] = [v[13], v[14], v[8], v[15]]
        u[i, :3] = [int(v[16]), int(v[17]), int(v[18])]
