"""The "CUDA torch backend" of the Krylov path: a ctypes binding of libcola_b200.so (C ABI in
include/cola_b200.h).  It plays the role cola/backends/torch_fns.py plays for the reference, but instead
of ~110 eager array functions it exposes the handful of fused kernels the loops need.

There is NO fallback: if the shared library is missing or an operand is not a contiguous CUDA tensor of
a supported dtype, calls raise.  torch is used only for device memory and streams.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcola_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "cola_b200.h")

_CTYPES = {
    "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "double": ctypes.c_double,
    "void": None,
}


def parse_header(path=HEADER_PATH):
    """Returns {name: (restype, [argtypes])} for every function declared in the C header, so the header stays
    the single source of truth for the ABI."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(cola_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "*" in ret:
            restype = ctypes.c_char_p if "char" in ret else ctypes.c_void_p
        else:
            restype = _CTYPES[ret.replace("const", "").strip()]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = a.replace("const", "").split()[0]
                    argtypes.append(_CTYPES[base])
        decls[name] = (restype, argtypes)
    return decls


class CgCtl(ctypes.Structure):
    """Mirror of cola_cg_ctl_t (device-resident; this host copy is only for reading it back)."""
    _fields_ = [("it", ctypes.c_int32), ("done", ctypes.c_int32), ("max_iters", ctypes.c_int32), ("k", ctypes.c_int32)]


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m cola_b200.build` "
                               "(cola_b200 has no CPU or eager-torch fallback)")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.decls = parse_header()
        for name, (restype, argtypes) in self.decls.items():
            fn = getattr(self.cdll, name)  # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = restype, argtypes

    def call(self, name, *args):
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            msg = self.cdll.cola_last_error()
            raise RuntimeError(f"{name} failed with status {rc}: {msg.decode() if msg else ''}")

    def launch_count(self):
        return int(self.cdll.cola_launch_count())


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


SUFFIX = {torch.float32: "f32", torch.float64: "f64"}


def sfx(dtype):
    try:
        return SUFFIX[dtype]
    except KeyError:
        raise TypeError(f"cola_b200 supports float32/float64 operators on CUDA, got {dtype}") from None


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what="operand"):
    """The single device gate of the package: every loop and every matmat passes its operands through here, and a
    CPU tensor raises (there is no CPU or eager-torch fallback)."""
    if not t.is_cuda:
        raise RuntimeError(f"cola_b200 is a CUDA-only path: {what} is on the CPU (no CPU fallback); "
                           "move the operator and the operand to a B200 with .to('cuda')")
    # kernels are enqueued on the CURRENT device's current stream (one process per GPU): an operand that lives on
    # another device of the same process would be dereferenced by the wrong GPU
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"{what} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                           "run the call under torch.cuda.device(...) (cola_b200 launches on the current device)")
    # the kernels do not record autograd: returning a result without grad_fn would silently cut a training graph
    if getattr(t, "requires_grad", False) and torch.is_grad_enabled():
        raise RuntimeError(f"{what} requires grad, but a plain matmat of cola_b200's native API is not differentiable (solves through CG "
                           "and the SLQ estimate are: cola_b200/autograd.py): detach it / use torch.no_grad(), or keep the reference's operators and call cola_b200.install(), under which "
                           "the reference's custom backward rules run with the solves on the kernels")


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (or NULL for None)."""
    if t is None:
        return None
    require_cuda(t, "a kernel operand")
    if not t.is_contiguous():
        raise RuntimeError("cola_b200 kernels need contiguous operands")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return ctypes.c_void_p(t.data_ptr())


def off_ptr(t, elem_offset):
    """Pointer to element `elem_offset` of contiguous CUDA tensor t."""
    require_cuda(t, "a kernel operand")
    if not t.is_contiguous():
        raise RuntimeError("cola_b200 kernels need contiguous operands")
    return ctypes.c_void_p(t.data_ptr() + elem_offset * t.element_size())


_PINNED = {}


def read_small(t):
    """Host copy of a small device tensor WITHOUT a DMA transfer (cola_publish_bytes): a one-block kernel writes the
    bytes into mapped pinned host memory in stream order, the host waits on an event and reads its own memory.
    `t.cpu()` would be a pageable cudaMemcpy, which queues behind whatever large transfer occupies the D2H copy engine
    (the polls of a solve whose previous 1 GiB result is still being copied out waited ~20 ms each)."""
    require_cuda(t, "a polled tensor")
    if not t.is_contiguous():
        t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    key = (t.device.index, torch.cuda.current_stream().cuda_stream)
    slot = _PINNED.get(key)
    if slot is None or slot[0].numel() < nbytes:
        slot = (torch.empty(max(4096, nbytes), dtype=torch.uint8).pin_memory(), torch.cuda.Event())
        _PINNED[key] = slot
    pin, ev = slot
    lib().call("cola_publish_bytes", ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(pin.data_ptr()), nbytes, stream_ptr())
    ev.record()
    ev.synchronize()
    return pin[:nbytes].clone().view(t.dtype).reshape(t.shape)


_PINNED_RING = {}


def publish_async(t):
    """Stream-ordered, DMA-free snapshot of a small device tensor (see read_small) that the host picks up LATER with
    publish_result: the Krylov loops enqueue the next batch of iterations before they look at the control block the previous
    batch left, so the device never idles through a host round trip.  A ring of four pinned slots per (device, stream)."""
    require_cuda(t, "a polled tensor")
    if not t.is_contiguous():
        t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    key = (t.device.index, torch.cuda.current_stream().cuda_stream)
    ring = _PINNED_RING.get(key)
    if ring is None:
        ring = {"slots": [(torch.empty(4096, dtype=torch.uint8).pin_memory(), torch.cuda.Event()) for _ in range(4)], "next": 0}
        _PINNED_RING[key] = ring
    if nbytes > 4096:
        raise ValueError("publish_async is for control blocks (<= 4 KB)")
    pin, ev = ring["slots"][ring["next"]]
    ring["next"] = (ring["next"] + 1) % 4
    lib().call("cola_publish_bytes", ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(pin.data_ptr()), nbytes, stream_ptr())
    ev.record()
    return (pin, ev, nbytes, t.dtype, tuple(t.shape))


def publish_result(token):
    pin, ev, nbytes, dtype, shape = token
    ev.synchronize()
    return pin[:nbytes].clone().view(dtype).reshape(shape)


def small_ints(values, device):
    """int32 device tensor of four values, written by a kernel from its arguments (cola_store_i32x4): no pageable
    H2D memcpy, which would queue behind a large transfer on the copy engine (see read_small)."""
    a, b, c, d = (int(v) for v in values)
    t = torch.empty(4, dtype=torch.int32, device=device)
    require_cuda(t, "a control block")
    lib().call("cola_store_i32x4", ptr(t), a, b, c, d, stream_ptr())
    return t


def scalar(dtype, x):
    return ctypes.c_float(x) if dtype == torch.float32 else ctypes.c_double(x)


# ------------------------------------------------------------------------------------------------
# thin tensor-level wrappers (one per C entry point family)
# ------------------------------------------------------------------------------------------------
def col_dots(X, Y, dots, gate=None):
    n, k = X.shape
    lib().call(f"cola_col_dots_{sfx(X.dtype)}", ptr(X), ptr(Y, X.dtype), n, k, k, ptr(dots, torch.float64), ptr(gate),
               stream_ptr())


def col_scale(X, Y, sq, take_sqrt, mode, a=1.0, gate=None):
    n, k = X.shape
    lib().call(f"cola_col_scale_{sfx(X.dtype)}", ptr(X), ptr(Y, X.dtype), n, k, k, ptr(sq, torch.float64),
               int(take_sqrt), int(mode), scalar(X.dtype, a), ptr(gate), stream_ptr())


def axpby(X, Y, a, b, gate=None):
    n, k = X.shape
    lib().call(f"cola_axpby_{sfx(X.dtype)}", ptr(X), ptr(Y, X.dtype), n, k, k, scalar(X.dtype, a), scalar(X.dtype, b),
               ptr(gate), stream_ptr())


def diag_matmat(X, Y, shift, diag, accumulate, dots=None, dots_row=None, gate=None):
    n, k = X.shape
    lib().call(f"cola_diag_matmat_{sfx(X.dtype)}", ptr(X), k, ptr(Y, X.dtype), k, n, k, scalar(X.dtype, shift),
               ptr(diag, X.dtype) if diag is not None else None, int(accumulate), ptr(dots), ptr(dots_row), ptr(gate),
               stream_ptr())


def csr_spmm(rowptr, colidx, vals, shape, nnz, max_row_nnz, X, Y, alpha=1.0, shift=0.0, diag=None, accumulate=False, dots=None,
             dots_row=None, gate=None):
    k = X.shape[1]
    dt = vals.dtype
    lib().call(f"cola_csr_spmm_{sfx(dt)}", ptr(rowptr, torch.int32), ptr(colidx, torch.int32), ptr(vals), shape[0],
               shape[1], nnz, max_row_nnz, ptr(X, dt), k, k, ptr(Y, dt), k, scalar(dt, alpha), scalar(dt, shift),
               ptr(diag, dt) if diag is not None else None, int(accumulate), ptr(dots), ptr(dots_row), ptr(gate),
               stream_ptr())


def csr_spmm_tiled(T, vals, shape, X, Y, alpha=1.0, shift=0.0, diag=None, accumulate=False, dots=None, dots_row=None,
                   gate=None):
    """Staged SpMM from the tile-local form `T` (cola_b200.csr_tiles.CsrTiles); `vals` = T.values(data)."""
    k = X.shape[1]
    dt = vals.dtype
    lib().call(f"cola_csr_spmm_tiled_{sfx(dt)}", ptr(T.rec, torch.int32), ptr(T.rp, torch.int32), ptr(T.idx, torch.int32),
               ptr(vals), shape[0], T.n_tiles, T.n_tiles2d, T.rows2d, T.stride, T.strip_rows, T.strips, T.cap_rows, T.cap_nz,
               ptr(X, dt), k, ptr(Y, dt), k, scalar(dt, alpha), scalar(dt, shift),
               ptr(diag, dt) if diag is not None else None, int(accumulate), ptr(dots), ptr(dots_row), ptr(gate),
               stream_ptr())


def mode_contract(M, d_out, d_in, pre, post, inp, out, alpha=1.0, shift=0.0, diag=None, epi_x=None, accumulate=False,
                  dots=None, dots_row=None, gate=None):
    """inp/out/diag/epi_x are ctypes pointers or tensors (tensors are converted)."""
    dt = M.dtype

    def p(x):
        return x if (x is None or isinstance(x, ctypes.c_void_p)) else ptr(x, dt)

    lib().call(f"cola_mode_contract_{sfx(dt)}", ptr(M), M.stride(0), d_out, d_in, pre, post, p(inp), p(out),
               scalar(dt, alpha), scalar(dt, shift), p(diag), p(epi_x), int(accumulate), ptr(dots), ptr(dots_row),
               ptr(gate), stream_ptr())


def mode_contract_tc_ok(M, pre, L, k, X):
    """True when cola_mode_contract_tc_f32 takes this mode: fp32, a square 64- or 128-wide factor, k % 32 == 0."""
    d = M.shape[0]
    return (M.dtype == torch.float32 and X.dtype == torch.float32 and M.shape[0] == M.shape[1] and
            bool(lib().cdll.cola_mode_contract_tc_supported(d, pre, L, k)))


def mode_contract_tc(M, pre, L, k, inp, out, alpha=1.0, shift=0.0, diag=None, epi_x=None, accumulate=False, dots=None,
                     dots_row=None, gate=None):
    """out[p, a, l, r] = alpha * sum_j M[a, j] inp[p, j, l, r] on the tensor cores (3xTF32); epilogue as mode_contract."""
    dt = torch.float32

    def p(x):   # tensors or ready-made pointers (row blocks of a BlockDiag operand)
        return x if (x is None or isinstance(x, ctypes.c_void_p)) else ptr(x, dt)

    lib().call("cola_mode_contract_tc_f32", ptr(M, dt), M.stride(0), M.shape[0], pre, L, k, p(inp), p(out),
               ctypes.c_float(alpha), ctypes.c_float(shift), p(diag), p(epi_x), int(accumulate), ptr(dots), ptr(dots_row),
               ptr(gate), stream_ptr())


def reorth_dots(V, j0, j1, W, C, gate=None):
    """V (n_vec, n, b) contiguous, W (n, b), C (n_vec, b) float64."""
    n, b = W.shape
    lib().call(f"cola_reorth_dots_{sfx(W.dtype)}", ptr(V, W.dtype), n * b, j0, j1, ptr(W), n, b, ptr(C, torch.float64),
               ptr(gate), stream_ptr())


def reorth_update(V, j0, j1, W, C, sign=-1.0, wnorm2=None, gate=None):
    n, b = W.shape
    lib().call(f"cola_reorth_update_{sfx(W.dtype)}", ptr(V, W.dtype), n * b, j0, j1, ptr(W), n, b,
               ptr(C, torch.float64), scalar(W.dtype, sign), ptr(wnorm2), ptr(gate), stream_ptr())


FUSED_REORTH = os.environ.get("COLA_NO_FUSED_REORTH", "") == ""


def reorth_update_dots(V, j0, j1, W, C1, C2, sign=-1.0, gate=None):
    """W += sign * V C1 and C2 += V^T W_new with one sweep over the basis.  Returns False (nothing launched) when
    the shape is outside the fused kernel's envelope; the caller then runs reorth_update + reorth_dots."""
    if not FUSED_REORTH:
        return False
    n, b = W.shape
    rc = getattr(lib().cdll, f"cola_reorth_update_dots_{sfx(W.dtype)}")(
        ptr(V, W.dtype), n * b, j0, j1, ptr(W), n, b, ptr(C1, torch.float64), scalar(W.dtype, sign),
        ptr(C2, torch.float64), ptr(gate), stream_ptr())
    if rc == -2:   # COLA_E_UNSUPPORTED
        return False
    if rc != 0:
        msg = lib().cdll.cola_last_error()
        raise RuntimeError(f"cola_reorth_update_dots failed with status {rc}: {msg.decode() if msg else ''}")
    return True


def lanczos_three_term(W, Vi, Vim1, alpha_acc, beta_prev_sq, gate=None):
    n, b = W.shape
    lib().call(f"cola_lanczos_three_term_{sfx(W.dtype)}", ptr(W), ptr(Vi, W.dtype),
               ptr(Vim1, W.dtype) if Vim1 is not None else None, n, b, ptr(alpha_acc, torch.float64),
               ptr(beta_prev_sq) if beta_prev_sq is not None else None, ptr(gate), stream_ptr())


def tridiag_eig_first_row(d, e, z, status=None, gate=None):
    """d, e, z: (m, b) float64 contiguous; see cola_tridiag_eig_first_row_f64."""
    m, b = d.shape
    lib().call("cola_tridiag_eig_first_row_f64", ptr(d, torch.float64), ptr(e, torch.float64), ptr(z, torch.float64), m, b,
               b, ptr(status, torch.int32) if status is not None else None, ptr(gate), stream_ptr())


def mgs_link(W, Qprev, hprev, Qcur, hcur, wnorm2=None, gate=None):
    n, b = W.shape
    lib().call(f"cola_mgs_link_{sfx(W.dtype)}", ptr(W), ptr(Qprev) if Qprev is not None else None,
               ptr(hprev) if hprev is not None else None, ptr(Qcur) if Qcur is not None else None,
               ptr(hcur) if hcur is not None else None, ptr(wnorm2) if wnorm2 is not None else None, n, b, ptr(gate),
               stream_ptr())


def mgs_chain(W, Q, n_links, H, wnorm2=None, gate=None):
    """All MGS links of one Arnoldi step in one cooperative launch: Q (m+1, n, b) basis, H (>= n_links, b) fp64 rows
    (zero on entry), W (n, b).  Returns False when the library declines (block too wide / launch refused): the caller
    then runs the mgs_link chain."""
    n, b = W.shape
    rc = getattr(lib().cdll, f"cola_mgs_chain_{sfx(W.dtype)}")(
        ptr(W), ptr(Q, W.dtype), Q.stride(0), n_links, ptr(H, torch.float64), H.stride(0),
        ptr(wnorm2, torch.float64) if wnorm2 is not None else None, n, b, ptr(gate), stream_ptr())
    if rc == -2:
        return False
    if rc != 0:
        msg = lib().cdll.cola_last_error()
        raise RuntimeError(f"cola_mgs_chain failed with status {rc}: {msg.decode() if msg else ''}")
    return True


# ---- parameter gradients (csrc/param_grad.cu) ------------------------------------------------------------
def sddmm_csr(rowptr, colidx, n_rows, G, V, alpha, out):
    """out[e] = alpha * <G[row(e), :], V[colidx[e], :]> over the CSR pattern; G, V (n, k) contiguous."""
    k = V.shape[1]
    dt = V.dtype
    lib().call(f"cola_sddmm_csr_{sfx(dt)}", ptr(rowptr, torch.int32), ptr(colidx, torch.int32), n_rows, ptr(G, dt), k,
               ptr(V, dt), k, k, scalar(dt, alpha), ptr(out, dt), 0, stream_ptr())


def row_dots(G, g_row0, V, v_row0, n, alpha, out, out_sq=None, accumulate=False):
    """out[i] (+)= alpha * <G[g_row0 + i, :], V[v_row0 + i, :]>, i < n; G, V (rows, k) contiguous.  With out_sq also
    out_sq[i] (+)= sum_c (G[.., c] V[.., c])^2."""
    k = V.shape[1]
    dt = V.dtype
    if n > 0:
        lib().call(f"cola_row_dots_{sfx(dt)}", off_ptr(G, g_row0 * k), k, off_ptr(V, v_row0 * k), k, n, k,
                   scalar(dt, alpha), ptr(out, dt), ptr(out_sq, dt) if out_sq is not None else None, int(accumulate),
                   stream_ptr())


def gram_nt(G, g_off, Z, z_off, d_g, d_z, pre, post, alpha, C):
    """C[a, j] += alpha * sum_{p, t} G[g_off + (p*d_g + a)*post + t] * Z[z_off + (p*d_z + j)*post + t]; C (d_g, d_z)
    float64, zeroed by the caller."""
    dt = Z.dtype
    lib().call(f"cola_gram_nt_{sfx(dt)}", off_ptr(G, g_off), off_ptr(Z, z_off), d_g, d_z, pre, post,
               ctypes.c_double(alpha), ptr(C, torch.float64), d_z, stream_ptr())
