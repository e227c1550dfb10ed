"""TEST INFRASTRUCTURE ONLY -- a stand-in for libcola_b200.so's kernels so that the HOST orchestration of the
Krylov loops (cola_b200/linalg/*.py, the operator plan compiler in cola_b200/ops.py) can be exercised in the
`-m "not gpu"` suite, in a container without a GPU.

`emulated_kernels()` replaces, for the duration of a `with` block, the tensor-level wrappers of
cola_b200/backend.py by torch-CPU statements of the semantics written next to each entry point in
include/cola_b200.h (same argument lists, same fp64 accumulators, same rounding points: coefficients are
rounded to the path dtype before use, reductions are taken in double), and opens the package's single device
gate (`backend.require_cuda`).  Nothing under cola_b200/ imports this module, the product has no switch that
reaches it, and outside the `with` block CPU tensors raise as always
(tests/test_abi_and_host.py::test_cpu_tensors_raise_no_fallback).  What it checks is loop logic -- stop rules,
indexing of the accumulator rows, the assembly of T / H / info, chunking, dispatch -- against the golden vectors;
the kernels themselves are only checked on the GPU (tests/test_gpu_parity.py, through the C ABI).  The CG loop
calls its four raw entry points (cola_cg_tol / advance / update_r / update_xp) through `lib().call`; those are
stated in `_CgEntryPoints` from the header and csrc/vec_kernels.cu, and CUDA-graph replay is switched off (the
eager batches of 16 gated iterations are the same launch sequence).
"""
import contextlib

import torch


def _open(gate):
    return gate is None or int(gate.reshape(-1)[0]) == 0


def _flat(t, numel):
    return t.reshape(-1)[:numel]


def _add_dots(dots, dots_row, k, X, Y):
    if dots is None:
        return
    r = 0 if dots_row is None else int(dots_row.reshape(-1)[0])
    acc = dots.reshape(-1)
    acc[r * k:(r + 1) * k] += (X.double() * Y.double()).sum(0)


def _epilogue(KX, X, Y, alpha, shift, diag, accumulate, dots, dots_row):
    """Y = alpha*KX + (shift + diag[i]) X (+Y); dots += <X, Y>."""
    res = alpha * KX
    if shift != 0.0:
        res = res + shift * X
    if diag is not None:
        res = res + _flat(diag, X.shape[0])[:, None] * X
    if accumulate:
        res = res + Y
    Y.copy_(res)
    _add_dots(dots, dots_row, X.shape[1], X, Y)


def col_dots(X, Y, dots, gate=None):
    if _open(gate):
        dots.reshape(-1)[:X.shape[1]] += (X.double() * Y.double()).sum(0)


def col_scale(X, Y, sq, take_sqrt, mode, a=1.0, gate=None):
    if not _open(gate):
        return
    s = sq.reshape(-1)[:X.shape[1]]
    s = (torch.sqrt(s) if take_sqrt else s).to(X.dtype)
    if mode == 0:
        Y.copy_(X * (a * s))
    elif mode == 1:
        Y.copy_(X / torch.where(s.abs() < 1e-40, torch.full_like(s, 1e-40), s))
    elif mode == 2:
        Y.copy_(X / s)
    else:
        Y.copy_(X / torch.clamp(s, min=a))


def axpby(X, Y, a, b, gate=None):
    if _open(gate):
        Y.copy_(a * X + b * Y if b != 0 else a * X)


def diag_matmat(X, Y, shift, diag, accumulate, dots=None, dots_row=None, gate=None):
    if _open(gate):
        _epilogue(torch.zeros_like(X), X, Y, 0.0, shift, diag, accumulate, dots, dots_row)


def csr_spmm(rowptr, colidx, vals, shape, nnz, max_row_nnz, X, Y, alpha=1.0, shift=0.0, diag=None, accumulate=False,
             dots=None, dots_row=None, gate=None):
    if not _open(gate):
        return
    # like the kernels, products are accumulated in double and rounded once to the path dtype
    S = torch.sparse_csr_tensor(rowptr.long(), colidx.long(), vals.double(), size=tuple(shape))
    KX = (S @ X.double()).to(X.dtype)
    if shift == 0.0 and diag is None and dots is None:
        Y.copy_(alpha * KX + Y if accumulate else alpha * KX)
    else:
        _epilogue(KX, X, Y, alpha, shift, diag, accumulate, dots, dots_row)


def csr_spmm_tiled(T, vals, shape, X, Y, alpha=1.0, shift=0.0, diag=None, accumulate=False, dots=None, dots_row=None,
                   gate=None):
    """What csrc/csr_tiled.cu computes, tile by tile FROM THE TILE-LOCAL FORM (records, runs, slots, local row pointers):
    a wrong record shows up here as a wrong product."""
    if not _open(gate):
        return
    n, k = shape[0], X.shape[1]
    KX = torch.full((n, k), float("nan"), dtype=torch.float64)
    RT = T.rows_per_tile
    lr_all = torch.arange(RT)
    for t in range(T.n_tiles):
        rec = T.rec[t].tolist()
        nzb, nr, nd = rec[0], rec[2], rec[3]
        if nr >= 0:
            runs = [(rec[8 + 2 * r], rec[9 + 2 * r] >> 16, rec[9 + 2 * r] & 0xFFFF) for r in range(nr)]
            staged = torch.zeros((max(nd, 1), k), dtype=torch.float64)
            for col0, slot0, ln in runs:
                staged[slot0:slot0 + ln] = X[col0:col0 + ln].double()
            assert sum(r[2] for r in runs) == nd and nd <= T.cap_rows
            src = staged
        else:
            src = X.double()
        rp = T.rp[t, :RT + 1].long()
        assert int(T.rp[t, RT + 3]) == nr
        cnt = rp[1:] - rp[:-1]
        nnz_t = int(rp[RT])
        assert nnz_t <= rec[1] <= T.cap_nz and rec[1] % 4 == 0 and nzb % 4 == 0
        lrow_e = torch.repeat_interleave(lr_all, cnt)
        e = slice(nzb, nzb + nnz_t)
        sel = T.idx[e].long()
        if nr >= 0:
            assert k * X.element_size() == T.row_bytes and bool((sel % T.row_bytes == 0).all())
            sel = sel // T.row_bytes
        contrib = vals[e].double()[:, None] * src[sel]
        Yt = torch.zeros((RT, k), dtype=torch.float64).index_add_(0, lrow_e, contrib)
        rows = torch.tensor([T.row_of(t, lr) for lr in range(RT)])
        ok = rows < n
        assert int(cnt[~ok].sum()) == 0
        if nr >= 0 and shape[0] == shape[1]:           # every row's own X row is staged (the fused epilogue reads it there)
            assert torch.equal(staged[(T.rp[t, RT + 4:2 * RT + 4].long() // T.row_bytes)[ok]], X[rows[ok]].double())
        KX[rows[ok]] = Yt[ok]
    assert not bool(torch.isnan(KX).any()), "a row belongs to no tile"
    KX = KX.to(X.dtype)
    if shift == 0.0 and diag is None and dots is None:
        Y.copy_(alpha * KX + Y if accumulate else alpha * KX)
    else:
        _epilogue(KX, X, Y, alpha, shift, diag, accumulate, dots, dots_row)


def mode_contract(M, d_out, d_in, pre, post, inp, out, alpha=1.0, shift=0.0, diag=None, epi_x=None, accumulate=False,
                  dots=None, dots_row=None, gate=None):
    if not _open(gate):
        return
    src = _flat(inp, pre * d_in * post).reshape(pre, d_in, post)
    dst = _flat(out, pre * d_out * post).reshape(pre * d_out, post)
    KX = torch.einsum("aj,pjq->paq", M[:d_out, :d_in].double(), src.double()).to(M.dtype).reshape(pre * d_out, post)
    if epi_x is None:
        dst.copy_(alpha * KX + dst if accumulate else alpha * KX)
        return
    X = _flat(epi_x, pre * d_out * post).reshape(pre * d_out, post)
    _epilogue(KX, X, dst, alpha, shift, diag, accumulate, dots, dots_row)


def reorth_dots(V, j0, j1, W, C, gate=None):
    if _open(gate) and j1 > j0:
        C[j0:j1] += (V[j0:j1].double() * W.double()[None]).sum(1)


def reorth_update(V, j0, j1, W, C, sign=-1.0, wnorm2=None, gate=None):
    if not _open(gate):
        return
    if j1 > j0:
        coef = C[j0:j1].to(W.dtype)
        W += sign * (coef[:, None, :] * V[j0:j1]).sum(0)
    if wnorm2 is not None:
        wnorm2.reshape(-1)[:W.shape[1]] += (W.double() ** 2).sum(0)


def lanczos_three_term(W, Vi, Vim1, alpha_acc, beta_prev_sq, gate=None):
    if not _open(gate):
        return
    upd = alpha_acc.to(W.dtype) * Vi
    if Vim1 is not None:
        upd = upd + torch.sqrt(beta_prev_sq).to(W.dtype) * Vim1
    W -= upd


def tridiag_eig_first_row(d, e, z, status=None, gate=None):
    if not _open(gate):
        return
    m, b = d.shape
    T = torch.diag_embed(d.T.contiguous())
    if m > 1:
        off = e[:m - 1].T.contiguous()
        T = T + torch.diag_embed(off, offset=1) + torch.diag_embed(off, offset=-1)
    lam, P = torch.linalg.eigh(T)
    d.copy_(lam.T)
    z.copy_(P[:, 0, :].T)
    if status is not None:
        status.zero_()


def mgs_link(W, Qprev, hprev, Qcur, hcur, wnorm2=None, gate=None):
    if not _open(gate):
        return
    b = W.shape[1]
    if Qprev is not None:
        W -= hprev.reshape(-1)[:b].to(W.dtype) * Qprev
    if Qcur is not None:
        hcur.reshape(-1)[:b] += (Qcur.double() * W.double()).sum(0)
    if wnorm2 is not None:
        wnorm2.reshape(-1)[:b] += (W.double() ** 2).sum(0)


def mgs_chain(W, Q, n_links, H, wnorm2=None, gate=None):
    if not _open(gate):
        return True
    mgs_link(W, None, None, Q[0], H[0])
    for j in range(1, n_links):
        mgs_link(W, Q[j - 1], H[j - 1], Q[j], H[j])
    mgs_link(W, Q[n_links - 1], H[n_links - 1], None, None, wnorm2=wnorm2)
    return True


def sddmm_csr(rowptr, colidx, n_rows, G, V, alpha, out):
    counts = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(n_rows), counts)
    out.copy_((alpha * (G.double()[rows] * V.double()[colidx.to(torch.int64)]).sum(1)).to(out.dtype))


def row_dots(G, g_row0, V, v_row0, n, alpha, out, out_sq=None, accumulate=False):
    pr = G.double()[g_row0:g_row0 + n] * V.double()[v_row0:v_row0 + n]
    res = (alpha * pr.sum(1)).to(out.dtype)
    out.copy_(out + res if accumulate else res)
    if out_sq is not None:
        res2 = (pr * pr).sum(1).to(out.dtype)
        out_sq.copy_(out_sq + res2 if accumulate else res2)


def gram_nt(G, g_off, Z, z_off, d_g, d_z, pre, post, alpha, C):
    g = G.reshape(-1)[g_off:g_off + pre * d_g * post].reshape(pre, d_g, post).double()
    z = Z.reshape(-1)[z_off:z_off + pre * d_z * post].reshape(pre, d_z, post).double()
    C += alpha * torch.einsum("pat,pjt->aj", g, z)


class _Fused:
    enabled = False


class _CgEntryPoints:
    """cola_cg_* of include/cola_b200.h on CPU tensors (be.ptr hands the tensors through).  ctl = int32
    [it, done, max_iters, k]; gamma / pAp are (rows, k) fp64 accumulators indexed by ctl.it."""
    cdll = None

    def call(self, name, *args):
        getattr(self, name.rsplit("_", 1)[0])(*args)

    def launch_count(self):
        return 0

    @staticmethod
    def _row(acc, it, k):
        return acc.reshape(-1)[it * k:(it + 1) * k]

    @classmethod
    def _alpha(cls, gamma, pap, it, k, dt):
        g = cls._row(gamma, it, k)
        tiny = torch.tensor(1e-40, dtype=dt)
        conv = torch.sqrt(g).to(dt) < tiny                       # has_converged, cg.py:144
        den = cls._row(pap, it, k).to(dt)
        den = torch.where(den.abs().double() < 1e-40, tiny, den)
        return torch.where(conv, torch.zeros((), dtype=dt), g.to(dt) / den), conv

    def cola_cg_tol(self, gamma0, tol, tol_eff, k, stream):
        dt = tol_eff.dtype
        t = torch.tensor(tol.value, dtype=dt)
        tol_eff.copy_(t * torch.sqrt(gamma0.reshape(-1)[:k]).to(dt) + t)

    def cola_cg_advance(self, ctl, gamma, tol_eff, increment, stream):
        if int(ctl[1]):
            return
        it, k = int(ctl[0]) + (1 if increment else 0), int(ctl[3])
        rs = torch.sqrt(self._row(gamma, it, k)).to(tol_eff.dtype)
        ctl[0] = it
        ctl[1] = 0 if (bool((rs > tol_eff).any()) and it < int(ctl[2])) else 1

    def cola_cg_update_r(self, R, AP, n, k, ld, ctl, gamma, pap, gamma_w, stream):
        if int(ctl[1]):
            return
        it = int(ctl[0])
        alpha, _ = self._alpha(gamma, pap, it, k, R.dtype)
        R -= alpha * AP
        self._row(gamma_w, it + 1, k).add_((R.double() ** 2).sum(0))

    def cola_cg_update_xp(self, X, R, P, n, k, ld, ctl, gamma, pap, stream):
        if int(ctl[1]):
            return
        it, dt = int(ctl[0]), X.dtype
        alpha, conv = self._alpha(gamma, pap, it, k, dt)
        g0 = self._row(gamma, it, k).to(dt)
        g0 = torch.where(g0.abs().double() < 1e-40, torch.tensor(1e-40, dtype=dt), g0)
        beta = torch.where(conv, torch.zeros((), dtype=dt), self._row(gamma, it + 1, k).to(dt) / g0)
        X += alpha * P
        P.copy_(R + beta * P)


def reorth_update_dots(V, j0, j1, W, C1, C2, sign=-1.0, gate=None):
    """False = "shape outside the fused kernel's envelope" (the caller's two-kernel path); with
    emulated_kernels(fused=True) the fused semantics are stated so both host branches are exercised."""
    if not _Fused.enabled:
        return False
    reorth_update(V, j0, j1, W, C1, sign=sign, gate=gate)
    reorth_dots(V, j0, j1, W, C2, gate=gate)
    return True


def mode_contract_tc(M, pre, L, k, inp, out, alpha=1.0, shift=0.0, diag=None, epi_x=None, accumulate=False, dots=None,
                     dots_row=None, gate=None):
    d = M.shape[0]
    mode_contract(M, d, d, pre, L * k, inp, out, alpha=alpha, shift=shift, diag=diag, epi_x=epi_x, accumulate=accumulate,
                  dots=dots, dots_row=dots_row, gate=gate)


_WRAPPERS = dict(mode_contract_tc=mode_contract_tc, mode_contract_tc_ok=lambda M, pre, L, k, X: False, col_dots=col_dots, col_scale=col_scale, axpby=axpby, diag_matmat=diag_matmat, csr_spmm=csr_spmm, csr_spmm_tiled=csr_spmm_tiled, mgs_chain=mgs_chain,
                 mode_contract=mode_contract, reorth_dots=reorth_dots, reorth_update=reorth_update,
                 reorth_update_dots=reorth_update_dots, lanczos_three_term=lanczos_three_term,
                 tridiag_eig_first_row=tridiag_eig_first_row, mgs_link=mgs_link,
                 sddmm_csr=sddmm_csr, row_dots=row_dots, gram_nt=gram_nt,
                 read_small=lambda t: t.detach().clone(), publish_async=lambda t: t.detach().clone(), publish_result=lambda tok: tok,
                 small_ints=lambda values, device: torch.tensor([int(v) for v in values], dtype=torch.int32))      # cola_publish_bytes: a host copy of a few device bytes


@contextlib.contextmanager
def emulated_kernels(fused=False):
    import importlib

    import cola_b200.backend as be
    import cola_b200.ops as ops
    import cola_b200.rng as rng
    cg = importlib.import_module("cola_b200.linalg.cg")
    saved = {name: getattr(be, name) for name in _WRAPPERS}
    saved_lib, saved_stream, saved_graph = be.lib, be.stream_ptr, cg.USE_CUDA_GRAPH
    saved_gate, saved_ptr, saved_off = be.require_cuda, be.ptr, be.off_ptr
    saved_tc, saved_mem, saved_probe = ops._KronCore._tc_ok, torch.cuda.mem_get_info, rng.PROBE_DEVICE
    try:
        for name, fn in _WRAPPERS.items():
            setattr(be, name, fn)
        be.require_cuda = lambda t, what="operand": None
        be.ptr = lambda t, dtype=None: t
        be.off_ptr = lambda t, elem_offset: t.reshape(-1)[elem_offset:]
        fake = _CgEntryPoints()
        be.lib, be.stream_ptr, cg.USE_CUDA_GRAPH = (lambda: fake), (lambda: None), False
        ops._KronCore._tc_ok = lambda self, X: False
        torch.cuda.mem_get_info = lambda device=None: (64 << 30, 64 << 30)
        rng.PROBE_DEVICE = "cpu"
        _Fused.enabled = fused
        yield
    finally:
        for name, fn in saved.items():
            setattr(be, name, fn)
        be.require_cuda, be.ptr, be.off_ptr = saved_gate, saved_ptr, saved_off
        be.lib, be.stream_ptr, cg.USE_CUDA_GRAPH = saved_lib, saved_stream, saved_graph
        ops._KronCore._tc_ok, torch.cuda.mem_get_info, rng.PROBE_DEVICE = saved_tc, saved_mem, saved_probe
        _Fused.enabled = False
