"""ctypes binding of libtrpx.so (include/trpx.h).  There is no CPU fallback: if the library is missing the
import of any compute entry point fails loudly."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TRPX_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libtrpx.so"))


class TrpxError(RuntimeError):
    pass


class Gpt2Config(ctypes.Structure):
    _fields_ = [("n_ctx", c_int), ("n_embd", c_int), ("n_layer", c_int), ("n_head", c_int), ("ln_eps", c_float)]


class MlpEmbConfig(ctypes.Structure):
    _fields_ = [("state_dim", c_int), ("hidden", c_int), ("n_embd", c_int), ("ln_eps", c_float)]


PP = POINTER(c_void_p)
# name -> (restype, argtypes); must list every symbol declared in include/trpx.h
SIGNATURES = {
    "trpx_last_error": (c_char_p, []),
    "trpx_version": (c_int, []),
    "trpx_device_info": (c_int, [c_int, POINTER(c_int), POINTER(c_int)]),
    "trpx_probe_fp32_fma": (c_int, [c_int, POINTER(c_float)]),
    "trpx_probe_tf32_mma": (c_int, [c_int, POINTER(c_float)]),
    "trpx_probe_umma_tf32": (c_int, [c_int, POINTER(c_float)]),
    "trpx_probe_tmem_ld": (c_int, [c_int, POINTER(c_float)]),
    "trpx_gpt2_train_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "trpx_gpt2_blob_size": (c_int, [c_void_p]),
    "trpx_gpt2_load_blob": (c_int, [c_void_p, c_void_p, c_void_p]),
    "trpx_probe_window_stream": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_float),
                                         POINTER(c_float)]),
    "trpx_gpt2_create": (c_int, [POINTER(Gpt2Config), c_int, POINTER(c_void_p)]),
    "trpx_gpt2_destroy": (c_int, [c_void_p]),
    "trpx_gpt2_num_weight_tensors": (c_int, [c_void_p]),
    "trpx_gpt2_load_weights": (c_int, [c_void_p, PP, c_int, c_void_p, c_void_p]),
    "trpx_gpt2_rollout": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int,
                                  c_void_p, c_void_p]),
    "trpx_gpt2_set_rollout_tuning": (c_int, [c_void_p, c_int, c_int, c_int, c_int]),
    "trpx_gpt2_last_launch": (c_int, [c_void_p] + [POINTER(c_int)] * 5),
    "trpx_gpt2_forward": (c_int, [c_void_p, c_void_p, PP, c_int, c_void_p, PP, PP, c_int, c_int, c_void_p]),
    "trpx_mlpemb_create": (c_int, [POINTER(MlpEmbConfig), c_int, POINTER(c_void_p)]),
    "trpx_mlpemb_destroy": (c_int, [c_void_p]),
    "trpx_mlpemb_load_weights": (c_int, [c_void_p, PP, c_int, c_void_p]),
    "trpx_mlpemb_embed": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "trpx_mlpemb_recover": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "trpx_mlpemb_set_recover_mode": (c_int, [c_void_p, c_int]),
    "trpx_mlpemb_recover_mse": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p]),
    "trpx_mlpemb_koopman": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "trpx_mlpemb_koopman_matrix": (c_int, [c_void_p, c_void_p, c_void_p]),
    "trpx_mlpemb_train_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float, c_float,
                                          c_void_p, c_void_p, c_void_p]),
    "trpx_mlpemb_num_params": (c_longlong, [c_void_p]),
    "trpx_clip_adam": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_float, c_float,
                               c_float, c_float, c_float, c_void_p, c_int, c_void_p]),
    "trpx_gs_create": (c_int, [c_int, c_float, c_int, POINTER(c_void_p)]),
    "trpx_gs_destroy": (c_int, [c_void_p]),
    "trpx_gs_num_weight_tensors": (c_int, [c_void_p]),
    "trpx_gs_load_weights": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_void_p]),
    "trpx_gs_embed": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "trpx_gs_recover": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "trpx_gs_decode_latent": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "trpx_gs_koopman": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "trpx_cyl_create": (c_int, [c_int, c_float, c_int, POINTER(c_void_p)]),
    "trpx_cyl_destroy": (c_int, [c_void_p]),
    "trpx_cyl_load_weights": (c_int, [c_void_p, PP, c_int, c_void_p]),
    "trpx_cyl_set_chunk": (c_int, [c_void_p, c_int]),
    "trpx_cyl_set_conv_mode": (c_int, [c_void_p, c_int]),
    "trpx_cyl_embed": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "trpx_cyl_recover": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "trpx_cyl_koopman": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "trpx_cyl_train_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float, c_float,
                                       c_void_p, c_void_p, c_void_p]),
    "trpx_cyl_num_params": (c_longlong, [c_void_p]),
}

_lib = None


def lib():
    """The loaded library (dlopen'ed on first use)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TrpxError(
                f"libtrpx.so not found at {LIB_PATH}: build it with `python transformer-physx_b200/build.py` "
                "(there is no CPU/PyTorch fallback for the rollout path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        raise TrpxError(f"libtrpx error {rc}: {lib().trpx_last_error().decode(errors='replace')}")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr_array(tensors):
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """Inputs must already live on a CUDA device in float32: the path has no CPU implementation."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise NotImplementedError(f"{name} is on {t.device}: trphysx_b200 runs on CUDA devices only (no CPU fallback)")
    if t.dtype != torch.float32:
        raise NotImplementedError(f"{name} has dtype {t.dtype}: the rollout path is float32 like the reference")
    return t.contiguous()


class HandleCache:
    """device index -> (native handle, WeightTracker) of ONE module, and the only owner of those handles.

    The native handles are destroyed by a `weakref.finalize` of this object, never by the module's `__del__`, and the
    cache itself is not copyable state: `copy.deepcopy(module)`, pickling (`torch.save(module)`) and
    `copy.copy(cache)` give the copy a fresh, empty cache whose handles are created lazily on first use, while
    shallow copies of the module's `__dict__` (`copy.copy(module)`, `nn.DataParallel` replicas) share THIS object, so
    a handle is freed exactly once, after the last module that can reach it is gone."""

    def __init__(self, destroy_symbol: str):
        import weakref
        self.destroy_symbol = destroy_symbol
        self._entries = {}
        self._finalizer = weakref.finalize(self, HandleCache._destroy_all, self._entries, destroy_symbol)

    @staticmethod
    def _destroy_all(entries, destroy_symbol):
        try:
            fn = getattr(lib(), destroy_symbol)
            for h, _ in entries.values():
                fn(h)
        except Exception:
            pass
        entries.clear()

    def get(self, idx):
        return self._entries.get(idx)

    def __setitem__(self, idx, entry):
        self._entries[idx] = entry

    def __getitem__(self, idx):
        return self._entries[idx]

    def __len__(self):
        return len(self._entries)

    def values(self):
        return self._entries.values()

    def invalidate(self):
        """Forget what was packed: the next call re-sends the module's weights to every handle."""
        for _, tracker in self._entries.values():
            tracker.key = None

    def __copy__(self):
        return HandleCache(self.destroy_symbol)

    def __deepcopy__(self, memo):
        return HandleCache(self.destroy_symbol)

    def __reduce__(self):
        return (HandleCache, (self.destroy_symbol,))


_weight_epoch = 0


def bump_weight_epoch():
    """Called by code that rewrites parameter storage from a native kernel (which torch's version counters
    cannot see), e.g. FusedClipAdam.step()."""
    global _weight_epoch
    _weight_epoch += 1


class WeightTracker:
    """Detects parameter changes (optimizer steps, load_state_dict, .to()) so packed weights are re-sent lazily."""

    def __init__(self):
        self.key = None

    def changed(self, tensors) -> bool:
        key = (_weight_epoch,) + tuple((t.data_ptr(), t._version, t.device) for t in tensors)
        if key != self.key:
            self.key = key
            return True
        return False
