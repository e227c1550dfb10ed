"""Tile-local form of a CSR pattern for the staged SpMM kernel (csrc/csr_tiled.cu, cola_csr_spmm_tiled_*).

The register-gather SpMM (csrc/csr_spmm.cu) fetches every gathered row of X from L2 once per CTA tile that touches it: for
a 5-point stencil on a D-wide grid that is 3 rows of X per output row (the row's own neighbourhood and the rows a grid
line above and below), 3.2 GB through L2 for a 1.07 GB block (ncu, profiles/r2_cg_cfg2_summary.md), and the kernel sits
at the ~8 TB/s the L2 fabric delivers.  The staged kernel makes the reuse explicit: a tile is S strips of R consecutive
rows, D rows apart (D = the pattern's dominant far diagonal, so the rows gathered across it are the neighbouring strips
of the SAME tile); the distinct columns a tile touches are brought into shared memory ONCE as a few contiguous runs of X
rows (bulk copies), and the non-zeros address them by slot.  This module computes, once per pattern, on the device:

  rec    (n_tiles, 32) int32   per tile: [nz_begin, nz_padded, n_runs (-1: irregular tile, gathers from global),
                               n_distinct, 0...] then per run (col0, slot0 << 16 | len) from word 8 on
  rp     (n_tiles, RP) int32   RP = 2 * rows per tile + 4: [0, RT] local row pointers of the tile's rows, [RT + 3] n_runs
                               again, [RT + 4, 2 RT + 4) the byte offset of each row's OWN staged X row (square operators; the
                               epilogue's operand comes from shared memory too)
  idx    (nz_padded_total,)    per non-zero: BYTE offset (slot * row_bytes) of its column's X row among the tile's staged
                               rows (irregular tiles: the column)
  perm   (nz_padded_total,)    position of each padded entry in the operator's `data` (-1 for padding): values are
                               gathered through it whenever `data` changes
A tile's non-zero range starts at a multiple of 4 entries (16-byte aligned bulk copies).
"""
import torch

STRIPS = 8            # S
MAX_RUNS = 12         # runs a record holds; tiles with more are "irregular"
REC_WORDS = 32


def far_diagonal(indices, row_indices, nnz, square):
    """Distance |col - row| of the pattern's dominant far diagonal (the grid width of a stencil matrix), or 0."""
    if nnz == 0 or not square:
        return 0
    step = max(1, nnz // (1 << 20))
    off = (indices[::step].to(torch.int64) - row_indices[::step].to(torch.int64)).abs()
    off = off[off >= 64]
    if off.numel() == 0:
        return 0
    vals, counts = torch.unique(off, return_counts=True)
    top = int(torch.argmax(counts))
    if int(counts[top]) * 8 >= (nnz + step - 1) // step:          # at least an eighth of all entries
        return int(vals[top])
    return 0


def banded_like(indices, row_indices, nnz):
    """Cheap screen before the tile form is built (sorts and uniques over every non-zero): do a handful of diagonals
    hold almost all entries?  Sampled; stencil and banded matrices pass, graphs without structure do not."""
    if nnz == 0:
        return False
    step = max(1, nnz // (1 << 20))
    off = indices[::step].to(torch.int64) - row_indices[::step].to(torch.int64)
    _, counts = torch.unique(off, return_counts=True)
    top = torch.sort(counts, descending=True).values[:32]
    return int(top.sum()) * 10 >= 9 * off.numel()


class CsrTiles:
    """See the module docstring.  `strip_rows` = largest R wanted (R becomes the largest divisor of D below it when the
    pattern has a far diagonal D), `cap_rows` = staged rows of X a tile may hold (shared-memory budget of one ring stage), `row_bytes` = bytes of one row of X (k * itemsize: the
    form is specific to it)."""

    def __init__(self, S, strip_rows, cap_rows, row_bytes):
        n_rows, n_cols = S.shape
        dev = S.indices.device
        R, NS = int(strip_rows), STRIPS
        D = far_diagonal(S.indices, S.row_indices, S.nnz, n_rows == n_cols)
        two_d = False
        if n_rows >= NS * D and D > 0:                     # strips D apart: the largest R <= strip_rows that divides D
            for r in range(min(R, D // 2), 7, -1):
                if D % r == 0:
                    R, two_d = r, True
                    break
        RT = R * NS
        self.strip_rows, self.strips, self.rows_per_tile, self.rp_stride = R, NS, RT, 2 * RT + 4
        self.row_bytes = int(row_bytes)
        self.stride = D if two_d else R
        n_blk = n_rows // (NS * D) if two_d else 0
        self.rows2d = n_blk * NS * D if two_d else 0
        self.tiles_per_blk = D // R if two_d else 1
        self.n_tiles2d = n_blk * self.tiles_per_blk
        n_tail = -(-(n_rows - self.rows2d) // RT)
        self.n_tiles = self.n_tiles2d + n_tail

        rows = S.row_indices.to(torch.int64)
        cols = S.indices.to(torch.int64)
        tile, lrow = self._tile_of(rows)
        order = torch.argsort(tile * RT + lrow, stable=True)            # by tile, then local row, original order within a row
        e_tile, e_lrow, e_col = tile[order], lrow[order], cols[order]
        t_nnz = torch.bincount(e_tile, minlength=self.n_tiles)
        nzp = (t_nnz + 3) // 4 * 4
        nz_ptr = torch.zeros(self.n_tiles + 1, dtype=torch.int64, device=dev)
        nz_ptr[1:] = torch.cumsum(nzp, 0)
        first_e = torch.zeros(self.n_tiles + 1, dtype=torch.int64, device=dev)
        first_e[1:] = torch.cumsum(t_nnz, 0)
        pos = nz_ptr[e_tile] + (torch.arange(e_tile.numel(), device=dev) - first_e[e_tile])   # padded position of every entry
        total = int(nz_ptr[-1])
        # local row pointers
        cnt = torch.bincount(e_tile * RT + e_lrow, minlength=self.n_tiles * RT).reshape(self.n_tiles, RT)
        rp = torch.zeros((self.n_tiles, self.rp_stride), dtype=torch.int32, device=dev)
        rp[:, 1:RT + 1] = torch.cumsum(cnt, 1).to(torch.int32)
        rp[:, RT + 1:RT + 3] = rp[:, RT:RT + 1]
        # distinct columns per tile -> slots and runs; a square operator also stages every row's own X row (the fused
        # epilogue's operand), whether or not the diagonal entry is stored
        keys = e_tile * n_cols + e_col
        if n_rows == n_cols:
            all_rows = torch.arange(n_rows, device=dev)
            s_tile, s_lrow = self._tile_of(all_rows)
            keys = torch.cat([keys, s_tile * n_cols + all_rows])
        ukey, inverse = torch.unique(keys, return_inverse=True)
        u_tile, u_col = ukey // n_cols, ukey % n_cols
        first_u = torch.searchsorted(u_tile, torch.arange(self.n_tiles + 1, device=dev))
        slot_u = torch.arange(ukey.numel(), device=dev) - first_u[u_tile]
        n_distinct = first_u[1:] - first_u[:-1]
        brk = torch.ones(ukey.numel(), dtype=torch.bool, device=dev)
        brk[1:] = (u_tile[1:] != u_tile[:-1]) | (u_col[1:] != u_col[:-1] + 1)
        r_start = brk.nonzero().reshape(-1)
        r_tile, r_col0, r_slot0 = u_tile[r_start], u_col[r_start], slot_u[r_start]
        r_len = torch.diff(r_start, append=torch.tensor([ukey.numel()], device=dev))
        n_runs = torch.bincount(r_tile, minlength=self.n_tiles)
        first_r = torch.zeros(self.n_tiles + 1, dtype=torch.int64, device=dev)
        first_r[1:] = torch.cumsum(n_runs, 0)
        regular = (n_runs <= MAX_RUNS) & (n_distinct <= cap_rows) & (n_distinct < 65536)
        rec = torch.zeros((self.n_tiles, REC_WORDS), dtype=torch.int64, device=dev)
        rec[:, 0] = nz_ptr[:-1]
        rec[:, 1] = nzp
        rec[:, 2] = torch.where(regular, n_runs, torch.full_like(n_runs, -1))
        rec[:, 3] = n_distinct
        keep = regular[r_tile]
        rr = (torch.arange(r_tile.numel(), device=dev) - first_r[r_tile])[keep]
        rec[r_tile[keep], 8 + 2 * rr] = r_col0[keep]
        rec[r_tile[keep], 9 + 2 * rr] = (r_slot0[keep] << 16) | r_len[keep]
        self.rec = rec.to(torch.int32).contiguous()
        idx = torch.zeros(total, dtype=torch.int32, device=dev)
        slot_e = slot_u[inverse[:e_tile.numel()]]
        rp[:, RT + 3] = rec[:, 2].to(torch.int32)
        if n_rows == n_cols:
            rp[s_tile, RT + 4 + s_lrow] = (slot_u[inverse[e_tile.numel():]] * self.row_bytes).to(torch.int32)
        self.rp = rp.contiguous()
        idx[pos] = torch.where(regular[e_tile], slot_e * self.row_bytes, e_col).to(torch.int32)
        self.idx = idx
        perm = torch.full((total, ), -1, dtype=torch.int64, device=dev)
        perm[pos] = order
        self.perm = perm
        self.cap_rows = int(n_distinct[regular].max()) if bool(regular.any()) else 0
        self.cap_nz = int(nzp.max()) if self.n_tiles else 0
        self.n_regular = int(regular.sum())
        self.mean_run = float(r_len[keep].double().mean()) if bool(keep.any()) else 0.0
        self._vals, self._vals_token = None, None

    def _tile_of(self, rows):
        R, NS, RT, D = self.strip_rows, STRIPS, self.rows_per_tile, self.stride
        tile = self.n_tiles2d + (rows - self.rows2d) // RT
        lrow = (rows - self.rows2d) % RT
        if self.rows2d > 0:
            in2d = rows < self.rows2d
            rem = rows % (NS * D)
            q = rem % D
            tile2 = (rows // (NS * D)) * self.tiles_per_blk + q // R
            lrow2 = (rem // D) * R + q % R
            tile = torch.where(in2d, tile2, tile)
            lrow = torch.where(in2d, lrow2, lrow)
        return tile, lrow

    def row_of(self, tile, lrow):
        """Global row of local row `lrow` of tile `tile` (what the kernel computes from the tile geometry)."""
        R, NS, RT, D = self.strip_rows, STRIPS, self.rows_per_tile, self.stride
        if tile < self.n_tiles2d:
            blk, c = divmod(tile, self.tiles_per_blk)
            return blk * NS * D + (lrow // R) * D + c * R + lrow % R
        return self.rows2d + (tile - self.n_tiles2d) * RT + lrow

    def values(self, data):
        """`data` in the padded tile order (zeros in the padding), refreshed when `data` was written in place or replaced."""
        token = (data.data_ptr(), data._version)
        if self._vals is None or self._vals_token != token:
            data = data.detach()
            v = torch.zeros(self.perm.numel(), dtype=data.dtype, device=data.device)
            m = self.perm >= 0
            v[m] = data[self.perm[m]]
            self._vals, self._vals_token = v, token
        return self._vals

    def stage_bytes(self, itemsize):
        """Bytes of one ring stage as csrc/csr_tiled.cu lays it out: X rows | offsets | values | row pointers, 128-byte aligned."""
        up = lambda x: -(-x // 128) * 128
        return up(self.cap_rows * self.row_bytes) + up(self.cap_nz * 4) + up(self.cap_nz * itemsize) + up(self.rp_stride * 4)

    def worthwhile(self):
        """The staged kernel pays when almost every tile is regular and its runs are long (stencil / banded patterns)."""
        return self.n_tiles > 0 and self.n_regular >= 0.9 * self.n_tiles and self.mean_run >= 8.0
