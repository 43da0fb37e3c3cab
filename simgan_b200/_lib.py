"""ctypes binding of the simgan_b200 C ABI (include/simgan_b200.h).

The product path has NO fallback: if the shared library is missing, ``lib()`` raises and every
hot-path call fails loudly.  Build it with ``python -m simgan_b200.build`` (or
``__graft_entry__.build()``).
"""
import ctypes as C
import os

from .build import LIB_PATH

_lib = None

c_float_p = C.POINTER(C.c_float)
c_void = C.c_void_p


class SgError(RuntimeError):
    pass


class PpoConfig(C.Structure):
    _fields_ = [("obs_dim", C.c_int), ("hidden", C.c_int), ("act_dim", C.c_int),
                ("T", C.c_int), ("N", C.c_int), ("ppo_epoch", C.c_int), ("num_mini_batch", C.c_int),
                ("mini_batch_size", C.c_int),
                ("clip_param", C.c_double), ("value_loss_coef", C.c_double), ("entropy_coef", C.c_double),
                ("max_grad_norm", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("adam_eps", C.c_double),
                ("use_clipped_value_loss", C.c_int), ("first_adam_step", C.c_int),
                ("row_begin", C.c_int), ("row_end", C.c_int), ("mode", C.c_int), ("dp_ctx", C.c_void_p)]


class DiscConfig(C.Structure):
    _fields_ = [("feat_dim", C.c_int), ("hidden", C.c_int), ("batch_size", C.c_int), ("n_steps", C.c_int),
                ("gp_lambda", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("adam_eps", C.c_double),
                ("first_adam_step", C.c_int), ("row_begin", C.c_int), ("row_end", C.c_int), ("mode", C.c_int),
                ("dp_ctx", C.c_void_p)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)

# name -> (restype, argtypes); every symbol include/simgan_b200.h declares
SIGNATURES = {
    "sg_last_error": (C.c_char_p, []),
    "sg_version": (C.c_int, []),
    "sg_launch_count": (C.c_longlong, []),
    "sg_device_sm_count": (C.c_int, []),
    "sg_policy_layout": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "sg_disc_layout": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "sg_compute_returns": (C.c_int, [c_void, c_void, c_void, c_void, c_void, c_void, C.c_int, C.c_int, C.c_double,
                                     C.c_double, C.c_int, C.c_int, c_void]),
    "sg_adv_stats_workspace_bytes": (C.c_int64, [C.c_int]),
    "sg_adv_stats": (C.c_int, [c_void, c_void, C.c_int, c_void, c_void, c_void]),
    "sg_gather_rows": (C.c_int, [C.POINTER(c_void), C.POINTER(c_void), C.POINTER(C.c_int), C.c_int, c_void, C.c_int,
                                 c_void]),
    "sg_copy_blocks": (C.c_int, [C.POINTER(c_void), C.POINTER(c_void), C.POINTER(C.c_int), C.c_int, c_void]),
    "sg_policy_forward": (C.c_int, [c_void, C.c_int, C.c_int, C.c_int, c_void, C.c_int, c_void, c_void, c_void, c_void,
                                    c_void, c_void, c_void]),
    "sg_rollout_stage_floats": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "sg_rollout_feed": (C.c_int, [c_void] + [C.c_int] * 7 + [c_void] * 13),
    "sg_ppo_workspace_bytes": (C.c_int64, [C.POINTER(PpoConfig)]),
    "sg_ppo_phase_cycles_offset": (C.c_int64, [C.POINTER(PpoConfig)]),
    "sg_ppo_uses_tensor_cores": (C.c_int, [C.POINTER(PpoConfig)]),
    "sg_ppo_update": (C.c_int, [C.POINTER(PpoConfig)] + [c_void] * 14 + [ALLREDUCE_FN, c_void, c_void]),
    "sg_split_layout": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "sg_split_forward": (C.c_int, [c_void, C.c_int, C.c_int, C.c_int, c_void, C.c_int, c_void, c_void, c_void, c_void,
                                   c_void, c_void, c_void]),
    "sg_split_ppo_workspace_bytes": (C.c_int64, [C.POINTER(PpoConfig)]),
    "sg_split_ppo_update": (C.c_int, [C.POINTER(PpoConfig)] + [c_void] * 15),
    "sg_dp_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sg_dp_local_handle": (C.c_int, [c_void, C.c_char_p]),
    "sg_dp_open_peers": (C.c_int, [c_void, C.c_char_p]),
    "sg_dp_capacity": (C.c_int, [c_void]),
    "sg_dp_destroy": (C.c_int, [c_void]),
    "sg_disc_workspace_bytes": (C.c_int64, [C.POINTER(DiscConfig)]),
    "sg_disc_phase_cycles_offset": (C.c_int64, [C.POINTER(DiscConfig)]),
    "sg_disc_update": (C.c_int, [C.POINTER(DiscConfig)] + [c_void] * 12 + [ALLREDUCE_FN, c_void, c_void]),
    "sg_disc_predict_reward": (C.c_int, [c_void, C.c_int, C.c_int, c_void, C.c_int, C.c_double, c_void, C.c_double,
                                         C.c_int, c_void, c_void, c_void]),
    "sg_relabel_workspace_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "sg_disc_relabel": (C.c_int, [c_void, C.c_int, C.c_int, c_void, c_void, c_void, C.c_int, C.c_int, C.c_double,
                                  C.c_double, c_void, C.c_int, c_void, c_void, c_void, c_void]),
    "sg_selftest_division": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_double, c_void, c_void]),
    "sg_selftest_mma": (C.c_int, [C.c_int] * 6 + [c_void] * 3 + [C.POINTER(C.c_int), c_void]),
    "sg_host_randperm_begin": (C.c_void_p, [c_void, C.c_int, C.c_int64, C.c_int, c_void, C.c_int, C.c_uint64]),
    "sg_host_randperm_prefix": (C.c_int, [c_void, C.c_int, C.c_int64, C.c_int64, c_void, c_void, C.POINTER(C.c_int)]),
    "sg_host_randperm_wait": (C.c_int, [c_void, C.c_int]),
    "sg_host_randperm_end": (C.c_int, [c_void, c_void, C.POINTER(C.c_int)]),
    "sg_relabel_normalize": (C.c_int, [c_void, c_void, c_void, C.c_int, C.c_int, C.c_double, c_void, C.c_int, c_void,
                                       c_void, c_void, c_void]),
}

NULL_ALLREDUCE = C.cast(None, ALLREDUCE_FN)


def lib():
    """Load (once) and return the C-ABI library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SgError("simgan_b200 CUDA library not built: %s is missing. Run `python -m simgan_b200.build` "
                          "(there is no CPU fallback for the PPO+GAIL hot path)." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def on_device(get_device):
    """Decorator for hot-path methods: run the body with the tensors' own CUDA device current.  Every sg_* entry point
    launches on the CURRENT device's current stream, so an object living on cuda:1 must not launch while cuda:0 is
    current (illegal addresses or silent peer access).  ``get_device(self, *args)`` returns a torch.device or None."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            import torch
            dev = get_device(*args, **kwargs)
            if dev is not None and dev.type == "cuda" and torch.cuda.is_available() and dev.index is not None \
                    and dev.index != torch.cuda.current_device():
                with torch.cuda.device(dev):
                    return fn(*args, **kwargs)
            return fn(*args, **kwargs)
        return wrapper
    return deco


def check(rc, what=""):
    if rc != 0:
        msg = lib().sg_last_error().decode("utf-8", "replace")
        raise SgError("%s failed (status %d): %s" % (what or "simgan_b200 call", rc, msg))


def ptr(t):
    """Device (or host) address of a torch tensor as a void*; None -> NULL."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


_readback = {}


def read_back(t):
    """Device tensor -> fresh host tensor through a cached pinned buffer: one async copy and one stream sync."""
    import torch
    key = (tuple(t.shape), t.dtype)
    buf = _readback.get(key)
    if buf is None:
        if len(_readback) > 64:
            _readback.clear()
        buf = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        _readback[key] = buf
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return buf.clone()


def policy_layout(obs_dim, hidden, act_dim):
    offs = (C.c_int * 13)()
    total = lib().sg_policy_layout(obs_dim, hidden, act_dim, offs)
    if total < 0:
        check(1, "sg_policy_layout")
    return list(offs), total


def split_layout(obs_dim, hidden, num_feet):
    offs = (C.c_int * 22)()
    total = lib().sg_split_layout(obs_dim, hidden, num_feet, offs)
    if total < 0:
        check(1, "sg_split_layout")
    return list(offs), total


def disc_layout(feat_dim, hidden):
    offs = (C.c_int * 6)()
    total = lib().sg_disc_layout(feat_dim, hidden, offs)
    if total < 0:
        check(1, "sg_disc_layout")
    return list(offs), total


class KernelTimer(object):
    """Optional CUDA-event timing of the hot-path launches, on the stream they are launched on.
    bench.py enables it to attribute the timed region to kernels (roofline.achieved); disabled (the
    default) it costs nothing."""

    def __init__(self):
        self.enabled = False
        self.records = []        # (name, start_event, end_event)

    def start(self, name):
        if not self.enabled:
            return None
        import torch
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        self.records.append((name, s, e))
        return e

    @staticmethod
    def stop(token):
        if token is not None:
            token.record()

    def drain(self):
        """{name: [ms, ...]} for every finished record; clears the list (caller synchronises first)."""
        out = {}
        for name, s, e in self.records:
            out.setdefault(name, []).append(s.elapsed_time(e))
        self.records = []
        return out


timer = KernelTimer()


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
