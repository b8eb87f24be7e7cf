"""Helpers for the -m gpu tests: device buffers come from torch (plumbing only), every call goes
through the C ABI of libprfdd_b200.so."""
import ctypes as C
import numpy as np
import torch


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def sync():
    torch.cuda.synchronize()


class WS:
    def __init__(self, lib):
        self.lib = lib
        self.h = C.c_void_p()
        assert lib.prfdd_reduce_ws_create(C.byref(self.h)) == 0

    def __del__(self):
        try:
            self.lib.prfdd_reduce_ws_destroy(self.h)
        except Exception:
            pass
