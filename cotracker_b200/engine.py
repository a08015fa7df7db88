"""ctypes binding of libct3_b200.so (C ABI: include/ct3_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every tensor is handed to the
library as a raw device pointer.  There is NO fallback: if the shared library is missing or a call
fails, a RuntimeError is raised (the product path never routes through the oracle or eager PyTorch).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CT3_B200_LIB", os.path.join(_HERE, "lib", "libct3_b200.so"))   # env override: A/B builds

LATENT, LEVELS, P, VOL, VOL_PAD = 128, 4, 49, 2401, 2432
HID, VIRT, XDIM, XDIM_PAD = 384, 64, 1110, 1152

_lib = None


class EngineError(RuntimeError):
    pass


def _declare(lib):
    c_int, c_size_t, c_void_p, c_char_p = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char_p
    i64p = ctypes.POINTER(ctypes.c_int64)
    intp = ctypes.POINTER(ctypes.c_int)
    sig = {
        "ct3_version": (c_int, []),
        "ct3_last_error": (c_char_p, []),
        "ct3_set_option": (c_int, [c_char_p, c_int]),
        "ct3_get_option": (c_int, [c_char_p, intp]),
        "ct3_precision_info": (c_int, [c_int, c_int, c_int, intp, intp, intp]),
        "ct3_volume_is_support_major": (c_int, [c_int, c_int, c_int, intp]),
        "ct3_num_weight_tensors": (c_int, []),
        "ct3_weight_name": (c_char_p, [c_int]),
        "ct3_packed_weights_bytes": (c_int, [ctypes.POINTER(c_size_t)]),
        "ct3_pack_weights": (c_int, [ctypes.POINTER(c_void_p), c_int, c_void_p, c_size_t, c_void_p]),
        "ct3_pyramid_layout": (c_int, [c_int, c_int, c_int, i64p, intp, intp, i64p]),
        "ct3_prepare_pyramid": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
        "ct3_sample_support": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
        "ct3_workspace_bytes": (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
        "ct3_update_loop": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
        "ct3_corr_sample": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
        "ct3_linear": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
        "ct3_split_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
        "ct3_split_rows_fp16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
        "ct3_linear_prec": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
        "ct3_updateformer": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
        "ct3_upsample_concat": (c_int, [ctypes.POINTER(c_void_p), intp, intp, intp, c_int, c_int, c_int, c_void_p, c_void_p]),
        "ct3_enc_tail_packed_bytes": (c_int, [ctypes.POINTER(c_size_t)]),
        "ct3_enc_tail_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
        "ct3_enc_tail_workspace_bytes": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
        "ct3_enc_tail": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
        "ct3_encoder_num_weight_tensors": (c_int, []),
        "ct3_encoder_weight_name": (c_char_p, [c_int]),
        "ct3_encoder_packed_bytes": (c_int, [ctypes.POINTER(c_size_t)]),
        "ct3_encoder_pack": (c_int, [ctypes.POINTER(c_void_p), c_int, c_void_p, c_size_t, c_void_p]),
        "ct3_encoder_workspace_bytes": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
        "ct3_encoder": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
        "ct3_profile_enable": (c_int, [c_int]),
        "ct3_profile_read": (c_int, [ctypes.POINTER(ctypes.c_double), intp, ctypes.POINTER(ctypes.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing -> loud
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTED_SYMBOLS = [
    "ct3_version", "ct3_last_error", "ct3_set_option", "ct3_get_option", "ct3_precision_info", "ct3_volume_is_support_major", "ct3_num_weight_tensors",
    "ct3_weight_name", "ct3_packed_weights_bytes", "ct3_pack_weights", "ct3_pyramid_layout",
    "ct3_prepare_pyramid", "ct3_sample_support", "ct3_workspace_bytes", "ct3_update_loop",
    "ct3_corr_sample", "ct3_linear", "ct3_linear_prec", "ct3_split_rows", "ct3_split_rows_fp16", "ct3_updateformer", "ct3_profile_enable", "ct3_profile_read",
    "ct3_encoder_num_weight_tensors", "ct3_encoder_weight_name", "ct3_encoder_packed_bytes", "ct3_encoder_pack",
    "ct3_encoder_workspace_bytes", "ct3_encoder",
    "ct3_upsample_concat", "ct3_enc_tail_packed_bytes", "ct3_enc_tail_pack", "ct3_enc_tail_workspace_bytes", "ct3_enc_tail",
]


def lib():
    """Load (once) and return the shared library; raises EngineError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C cotracker_b200/csrc`). There is no CPU/eager fallback.")
        try:
            handle = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise EngineError(f"cannot load {LIB_PATH}: {e}") from e
        _declare(handle)
        _lib = handle
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = lib().ct3_last_error().decode("utf-8", "replace")
        raise EngineError(f"{what} failed (code {rc}): {msg}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise EngineError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise EngineError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise EngineError(f"{name} must be contiguous")
    return t


def set_option(name: str, value: int):
    _check(lib().ct3_set_option(name.encode(), int(value)), f"ct3_set_option({name})")


def get_option(name: str) -> int:
    v = ctypes.c_int(0)
    _check(lib().ct3_get_option(name.encode(), ctypes.byref(v)), f"ct3_get_option({name})")
    return v.value


def precision_info(T: int = 16, H4: int = 96, W4: int = 128):
    """-> (corr_products, fc1_products, volume_bytes_per_element) in effect for this thread's options."""
    a, b, c = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    _check(lib().ct3_precision_info(T, H4, W4, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)), "ct3_precision_info")
    return a.value, b.value, c.value


def precision_summary(T: int = 16, H4: int = 96, W4: int = 128) -> dict:
    """What bench.py prints as `dtype`: the arithmetic each GEMM group computes in (no precision claim beyond it)."""
    corr, fc1, vb = precision_info(T, H4, W4)
    name = {3: "x3 (split x split: hi*hi+lo*hi+hi*lo)", 2: "x2 (fp16 plane x split fp16)", 1: "x1 (single fp16 product)"}
    return {
        "dtype": (f"transformer + corr_mlp.fc2 + input_transform: bf16x3; correlation einsum: "
                  f"{'bf16' if corr == 3 else 'fp16'}{name[corr][:2]}; corr_mlp.fc1: {'bf16' if fc1 == 3 else 'fp16'}"
                  f"{name[fc1][:2]}; fp32 accumulate, fp32 softmax/LayerNorm/GELU"),
        "products": f"3 (bf16 split) except correlation einsum {corr} and corr_mlp.fc1 {fc1}",
        "corr_products": corr, "fc1_products": fc1, "volume_bytes_per_element": vb,
    }


def weight_names() -> List[str]:
    L = lib()
    return [L.ct3_weight_name(i).decode() for i in range(L.ct3_num_weight_tensors())]


def packed_weights_bytes() -> int:
    n = ctypes.c_size_t(0)
    _check(lib().ct3_packed_weights_bytes(ctypes.byref(n)), "ct3_packed_weights_bytes")
    return n.value


def pack_weights(state: dict, device) -> torch.Tensor:
    """state: mapping state-dict key -> tensor (any device); returns the packed device buffer."""
    names = weight_names()
    tensors = []
    for k in names:
        if k not in state:
            raise EngineError(f"missing weight '{k}'")
        tensors.append(state[k].detach().to(device=device, dtype=torch.float32).contiguous())
    arr = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    nbytes = packed_weights_bytes()
    packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _check(lib().ct3_pack_weights(arr, len(tensors), _ptr(packed), nbytes, _stream(device)), "ct3_pack_weights")
        torch.cuda.current_stream(device).synchronize()  # `tensors` may be temporaries
    return packed


def pyramid_layout(T: int, H4: int, W4: int):
    off = (ctypes.c_int64 * 4)()
    h = (ctypes.c_int * 4)()
    w = (ctypes.c_int * 4)()
    tot = ctypes.c_int64(0)
    _check(lib().ct3_pyramid_layout(T, H4, W4, off, h, w, ctypes.byref(tot)), "ct3_pyramid_layout")
    return list(off), list(h), list(w), tot.value


def prepare_pyramid(fmaps: torch.Tensor) -> torch.Tensor:
    """fmaps [T,128,H4,W4] fp32 (raw fnet output) -> flat channels-last normalised 4-level pyramid."""
    _req(fmaps, torch.float32, "fmaps")
    T, C, H4, W4 = fmaps.shape
    if C != LATENT:
        raise EngineError("fmaps must have 128 channels")
    *_, total = pyramid_layout(T, H4, W4)
    pyr = torch.empty(total, dtype=torch.float32, device=fmaps.device)
    with torch.cuda.device(fmaps.device):
        _check(lib().ct3_prepare_pyramid(_ptr(fmaps), T, H4, W4, _ptr(pyr), _stream(fmaps.device)), "ct3_prepare_pyramid")
    return pyr


def pyramid_levels(pyr: torch.Tensor, T: int, H4: int, W4: int) -> List[torch.Tensor]:
    """Views [T,Hl,Wl,128] into the flat pyramid (tests / debugging)."""
    off, h, w, _ = pyramid_layout(T, H4, W4)
    return [pyr[off[l]: off[l] + T * h[l] * w[l] * LATENT].view(T, h[l], w[l], LATENT) for l in range(LEVELS)]


def sample_support(pyr, T, H4, W4, qframes, qcoords, support=None, accumulate_mask=None) -> torch.Tensor:
    _req(pyr, torch.float32, "pyr")
    _req(qframes, torch.int32, "queried_frames")
    _req(qcoords, torch.float32, "queried_coords")
    N = qframes.shape[0]
    if support is None:
        support = torch.zeros(LEVELS, P, N, LATENT, dtype=torch.float32, device=pyr.device)
    _req(support, torch.float32, "support")
    if accumulate_mask is not None:
        _req(accumulate_mask, torch.uint8, "accumulate_mask")
    with torch.cuda.device(pyr.device):
        _check(lib().ct3_sample_support(_ptr(pyr), T, H4, W4, _ptr(qframes), _ptr(qcoords), N, _ptr(accumulate_mask),
                                        _ptr(support), _stream(pyr.device)), "ct3_sample_support")
    return support


def workspace_bytes(T: int, N: int, H4: int = 0, W4: int = 0) -> int:
    """Scratch of ct3_update_loop for T frames of H4 x W4 feature maps and N tracks (H4 = W4 = 0: updateformer only)."""
    n = ctypes.c_size_t(0)
    _check(lib().ct3_workspace_bytes(T, N, H4, W4, ctypes.byref(n)), "ct3_workspace_bytes")
    return n.value


class WorkspaceCache:
    """Caller-owned scratch, grown on demand (the library never allocates)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    def get(self, T: int, N: int, device, H4: int = 0, W4: int = 0) -> torch.Tensor:
        need = workspace_bytes(T, N, H4, W4)
        if self.buf is None or self.buf.numel() < need or self.buf.device != torch.device(device):
            self.buf = None
            self.buf = torch.empty(need, dtype=torch.uint8, device=device)
        return self.buf


def update_loop(packed, pyr, H4, W4, support, track_valid, coords, vis, conf, time_emb, iters, workspace):
    """In-place refinement of coords [T,N,2], vis [T,N], conf [T,N] (fp32, feature-grid units / logits)."""
    _req(coords, torch.float32, "coords"); _req(vis, torch.float32, "vis"); _req(conf, torch.float32, "conf")
    _req(pyr, torch.float32, "pyr"); _req(support, torch.float32, "support"); _req(time_emb, torch.float32, "time_emb")
    T, N, _ = coords.shape
    if time_emb.shape != (T, XDIM):
        raise EngineError(f"time_emb must be [{T},{XDIM}]")
    if track_valid is not None:
        _req(track_valid, torch.uint8, "track_valid")
    with torch.cuda.device(coords.device):
        _check(lib().ct3_update_loop(_ptr(packed), _ptr(pyr), H4, W4, _ptr(support), _ptr(track_valid), _ptr(coords),
                                     _ptr(vis), _ptr(conf), _ptr(time_emb), T, N, int(iters), _ptr(workspace),
                                     workspace.numel(), _stream(coords.device)), "ct3_update_loop")


# ---- stage-level wrappers (tests, profiles) -----------------------------------------------------------
def corr_sample(pyr, H4, W4, support, track_valid, coords, scratch: bool = True) -> torch.Tensor:
    """-> fp32 correlation volume [N, T, 4, 2401] reconstructed from the split-bf16 device layout.
    scratch=False withholds the split-pyramid scratch, i.e. selects the sample-then-correlate kernel."""
    T, N, _ = coords.shape
    vol = torch.zeros(N * T * LEVELS, 2 * VOL_PAD, dtype=torch.bfloat16, device=coords.device)
    scr = torch.empty(pyr.numel() * 4, dtype=torch.uint8, device=coords.device) if scratch else None
    # a single fp16 plane [rows, 2432] when the correlate-then-interpolate kernel runs with prec.fc1 < 3
    vb = precision_info(T, H4, W4)[2] if (scratch and get_option("corr") in (0, 3)) else 4
    with torch.cuda.device(coords.device):
        _check(lib().ct3_corr_sample(_ptr(pyr), H4, W4, _ptr(support), _ptr(track_valid), _ptr(coords), T, N, _ptr(vol),
                                     _ptr(scr), scr.numel() if scratch else 0, _stream(coords.device)),
               "ct3_corr_sample")
    if vb == 2:
        full = vol.reshape(-1).view(torch.float16)[:N * T * LEVELS * VOL_PAD].reshape(-1, VOL_PAD).float()
    else:
        v = vol.float()
        full = v[:, :VOL_PAD] + v[:, VOL_PAD:]
    assert bool((full[:, VOL:] == 0).all()), "K padding of the correlation volume must be zero"
    full = full[:, :VOL]
    flag = ctypes.c_int(0)
    _check(lib().ct3_volume_is_support_major(T, H4, W4, ctypes.byref(flag)), "ct3_volume_is_support_major")
    if scratch and flag.value:   # corr_tc3.cu: rows are [k][a*7+b]; hand back the reference order [(a*7+b)][k]
        full = full.reshape(-1, P, P).transpose(1, 2).reshape(-1, VOL)
    return full.reshape(N, T, LEVELS, VOL)


def split_rows(x: torch.Tensor, Kpad: int, fp16: bool = False) -> torch.Tensor:
    _req(x, torch.float32, "x")
    rows, K = x.shape
    out = torch.empty(rows, 2 * Kpad, dtype=torch.bfloat16, device=x.device)   # 16-bit planes (bf16 or fp16 bits)
    fn = lib().ct3_split_rows_fp16 if fp16 else lib().ct3_split_rows
    with torch.cuda.device(x.device):
        _check(fn(_ptr(x), rows, K, Kpad, _ptr(out), _stream(x.device)), "ct3_split_rows")
    return out


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: int = 0, products: int = 3,
           fp16: bool = False) -> torch.Tensor:
    """Y = act(x w^T + b) through the tcgen05 GEMM engine; x [M,K], w [Nout,K] fp32.  products / fp16: the
    precision switches (3 = split x split, 2 = x_hi x split w, 1 = x_hi x w_hi; bf16 or fp16 planes)."""
    M, K = x.shape
    Nout = w.shape[0]
    Kpad = (K + 63) // 64 * 64
    xs, ws = split_rows(x.contiguous(), Kpad, fp16), split_rows(w.contiguous(), Kpad, fp16)
    y = torch.empty(M, Nout, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib().ct3_linear_prec(_ptr(xs), _ptr(ws), _ptr(bias), M, Nout, Kpad, act, products, 1 if fp16 else 0,
                                     _ptr(y), _stream(x.device)), "ct3_linear_prec")
    return y


def updateformer(packed, x: torch.Tensor, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [N,T,1110] fp32 (reference column order, time embedding added) -> delta [N,T,4]."""
    _req(x, torch.float32, "x")
    N, T, D = x.shape
    if D != XDIM:
        raise EngineError("x must be [N,T,1110]")
    if workspace is None:
        workspace = torch.empty(workspace_bytes(T, N), dtype=torch.uint8, device=x.device)
    delta = torch.empty(N, T, 4, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib().ct3_updateformer(_ptr(packed), _ptr(x), T, N, _ptr(delta), _ptr(workspace), workspace.numel(),
                                      _stream(x.device)), "ct3_updateformer")
    return delta


PROFILE_CATEGORIES = ["corr_sample", "gemm", "attention", "layernorm", "misc", "encoder", "qkv_time_attention"]


def profile_enable(on: bool):
    _check(lib().ct3_profile_enable(1 if on else 0), "ct3_profile_enable")


def profile_read():
    """-> ({category: ms}, {category: launches}, gemm_flops) accumulated since profile_enable(True)."""
    ms = (ctypes.c_double * len(PROFILE_CATEGORIES))()
    n = (ctypes.c_int * len(PROFILE_CATEGORIES))()
    fl = ctypes.c_double(0)
    _check(lib().ct3_profile_read(ms, n, ctypes.byref(fl)), "ct3_profile_read")
    return dict(zip(PROFILE_CATEGORIES, list(ms))), dict(zip(PROFILE_CATEGORIES, list(n))), fl.value


# ---- encoder tail -------------------------------------------------------------------------------------
def enc_tail_pack(conv2_w, conv2_b, conv3_w, conv3_b, device) -> torch.Tensor:
    n = ctypes.c_size_t(0)
    _check(lib().ct3_enc_tail_packed_bytes(ctypes.byref(n)), "ct3_enc_tail_packed_bytes")
    ts = [x.detach().to(device=device, dtype=torch.float32).contiguous() for x in (conv2_w, conv2_b, conv3_w, conv3_b)]
    packed = torch.empty(n.value, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _check(lib().ct3_enc_tail_pack(_ptr(ts[0]), _ptr(ts[1]), _ptr(ts[2]), _ptr(ts[3]), _ptr(packed), n.value,
                                       _stream(device)), "ct3_enc_tail_pack")
        torch.cuda.current_stream(device).synchronize()
    return packed


def enc_tail_workspace_bytes(T: int, H4: int, W4: int) -> int:
    n = ctypes.c_size_t(0)
    _check(lib().ct3_enc_tail_workspace_bytes(T, H4, W4, ctypes.byref(n)), "ct3_enc_tail_workspace_bytes")
    return n.value


def enc_tail(packed, cat: torch.Tensor, workspace: torch.Tensor) -> torch.Tensor:
    """cat [T,416,H4,W4] fp32 -> flat channels-last normalised pyramid (same layout as prepare_pyramid)."""
    _req(cat, torch.float32, "cat")
    T, C, H4, W4 = cat.shape
    if C != 416:
        raise EngineError("cat must have 416 channels")
    *_, total = pyramid_layout(T, H4, W4)
    pyr = torch.empty(total, dtype=torch.float32, device=cat.device)
    with torch.cuda.device(cat.device):
        _check(lib().ct3_enc_tail(_ptr(packed), _ptr(cat), T, H4, W4, _ptr(pyr), _ptr(workspace), workspace.numel(),
                                  _stream(cat.device)), "ct3_enc_tail")
    return pyr


# ---- whole encoder (enc_front.cu + GEMM engine) -----------------------------------------------------------
def encoder_weight_names() -> List[str]:
    L = lib()
    return [L.ct3_encoder_weight_name(i).decode() for i in range(L.ct3_encoder_num_weight_tensors())]


def encoder_pack(state: dict, device) -> torch.Tensor:
    """state: mapping `fnet` state-dict key (without the `fnet.` prefix) -> tensor; returns the packed device buffer."""
    names = encoder_weight_names()
    ts = [state[k].detach().to(device=device, dtype=torch.float32).contiguous() for k in names]
    n = ctypes.c_size_t(0)
    _check(lib().ct3_encoder_packed_bytes(ctypes.byref(n)), "ct3_encoder_packed_bytes")
    packed = torch.empty(n.value, dtype=torch.uint8, device=device)
    ptrs = (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    with torch.cuda.device(device):
        _check(lib().ct3_encoder_pack(ptrs, len(ts), _ptr(packed), n.value, _stream(device)), "ct3_encoder_pack")
        torch.cuda.current_stream(device).synchronize()   # `ts` may be temporaries
    return packed


def encoder_workspace_bytes(T: int, H: int, W: int) -> int:
    n = ctypes.c_size_t(0)
    _check(lib().ct3_encoder_workspace_bytes(T, H, W, ctypes.byref(n)), "ct3_encoder_workspace_bytes")
    return n.value


def encoder(packed: torch.Tensor, frames: torch.Tensor, workspace: torch.Tensor) -> torch.Tensor:
    """frames [T,3,H,W] fp32 in [-1,1] -> flat channels-last L2-normalised 4-level pyramid (stride 4)."""
    _req(frames, torch.float32, "frames")
    T, C, H, W = frames.shape
    if C != 3:
        raise EngineError("frames must be [T,3,H,W]")
    *_, total = pyramid_layout(T, H // 4, W // 4)
    pyr = torch.empty(total, dtype=torch.float32, device=frames.device)
    with torch.cuda.device(frames.device):
        _check(lib().ct3_encoder(_ptr(packed), _ptr(frames), T, H, W, _ptr(pyr), _ptr(workspace), workspace.numel(),
                                 _stream(frames.device)), "ct3_encoder")
    return pyr


def slice_pyramid(pyr: torch.Tensor, T: int, H4: int, W4: int, t0: int, S: int) -> torch.Tensor:
    """Frames [t0, t0+S) of every level as a new flat pyramid (sliding-window mode)."""
    off, h, w, _ = pyramid_layout(T, H4, W4)
    parts = []
    for l in range(LEVELS):
        per = h[l] * w[l] * LATENT
        parts.append(pyr[off[l] + t0 * per: off[l] + (t0 + S) * per])
    return torch.cat(parts)


def concat_pyramid_frames(pyr_a: torch.Tensor, Ta: int, a0: int, pyr_b: torch.Tensor, Tb: int, H4: int, W4: int):
    """Flat pyramid holding frames [a0, Ta) of `pyr_a` followed by all Tb frames of `pyr_b` (online feature reuse)."""
    off_a, h, w, _ = pyramid_layout(Ta, H4, W4)
    off_b, _, _, _ = pyramid_layout(Tb, H4, W4)
    parts = []
    for l in range(LEVELS):
        per = h[l] * w[l] * LATENT
        parts.append(pyr_a[off_a[l] + a0 * per: off_a[l] + Ta * per])
        parts.append(pyr_b[off_b[l]: off_b[l] + Tb * per])
    return torch.cat(parts)


def upsample_concat(feats: Sequence[torch.Tensor], H: int, W: int) -> torch.Tensor:
    """4 stage outputs [T,Cs,Hs,Ws] -> bilinear(align_corners=True) to HxW, concatenated on channels [T,sum Cs,H,W]."""
    if len(feats) != 4:
        raise EngineError("upsample_concat expects 4 stage tensors")
    fs = [_req(f, torch.float32, "stage feature") for f in feats]
    T = fs[0].shape[0]
    out = torch.empty(T, sum(f.shape[1] for f in fs), H, W, dtype=torch.float32, device=fs[0].device)
    src = (ctypes.c_void_p * 4)(*[f.data_ptr() for f in fs])
    ch = (ctypes.c_int * 4)(*[f.shape[1] for f in fs])
    hh = (ctypes.c_int * 4)(*[f.shape[2] for f in fs])
    ww = (ctypes.c_int * 4)(*[f.shape[3] for f in fs])
    with torch.cuda.device(out.device):
        _check(lib().ct3_upsample_concat(src, ch, hh, ww, T, H, W, _ptr(out), _stream(out.device)), "ct3_upsample_concat")
    return out
