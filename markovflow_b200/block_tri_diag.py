"""Block-tridiagonal operators with the API of ``markovflow/block_tri_diag.py``, executed by the
hand-written sm_100a kernels behind the C ABI (``include/markovflow_b200.h``).

Same class names, constructor signatures, properties, method names, shapes and broadcasting rules
as the reference (``block_tri_diag.py:37-592``).  Differences, by design:

* tensors are CUDA tensors (torch, or anything DLPack-exportable such as TF GPU tensors);
* the kernels consume the block layout directly -- the band layout the reference round-trips
  through (``_convert_to_band`` :206-237, ``_banded_to_block_tri`` :549-592) is never materialised
  (``as_band`` exists for API compatibility and debugging only);
* shape errors raise ``ValueError`` (reference: ``tf.errors.InvalidArgumentError``), a non-positive
  pivot raises :class:`CholeskyError` (reference: TF error "Banded Cholesky decomposition failure").
"""
from __future__ import annotations

import abc
import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import CholeskyError, check, current_stream, dtype_code, i64, ptr
from .config import check_numerics
from .autograd import CholeskyFn, SolveFn, needs_grad
from .interop import framework_of, as_torch, boundary, require_cuda


def _prod(shape) -> int:
    return int(math.prod(tuple(shape)))


def _raise_if_failed(info: torch.Tensor, what: str) -> None:
    if not check_numerics():
        return
    if not bool(torch.any(info)):  # one reduction + one 1-byte read on the success path
        return
    bad = torch.nonzero(info)
    if bad.numel():
        b = int(bad[0, 0])
        raise CholeskyError(
            f"Banded Cholesky decomposition failure in {what}: chain {b}, block {int(info[b])} "
            "(1-based) has a non-positive pivot"
        )


class BlockTriDiagonal(abc.ABC):
    """Abstract block-tridiagonal matrix (reference ``block_tri_diag.py:37-288``)."""

    def __init__(self, diagonal, symmetric: bool, sub_diagonal=None) -> None:
        self._fw = framework_of(diagonal, sub_diagonal)  # results come back in the caller's framework
        diagonal = as_torch(diagonal)
        if diagonal.dim() < 3:
            raise ValueError("diagonal must have shape [..., outer_dim, inner_dim, inner_dim]")
        if diagonal.shape[-1] != diagonal.shape[-2]:
            raise ValueError("Last two dimensions of the block diagonal must match.")
        self._diag = diagonal
        if sub_diagonal is not None:
            sub_diagonal = as_torch(sub_diagonal)
            if self.outer_dim <= 1:
                raise ValueError("There is no sub-diagonal with outer dimension of one.")
            want = tuple(self.batch_shape) + (self.outer_dim - 1, self.inner_dim, self.inner_dim)
            if tuple(sub_diagonal.shape) != want:
                raise ValueError(
                    f"Sub_diagonal has shape {tuple(sub_diagonal.shape)} but must have shape: {want}"
                )
            if sub_diagonal.dtype != diagonal.dtype or sub_diagonal.device != diagonal.device:
                raise ValueError("diagonal and sub_diagonal must share dtype and device")
        self._sub_diag = sub_diagonal
        self._symmetric = symmetric

    @property
    def _dev(self) -> torch.device:
        return self._diag.device

    # -- shape properties (reference :100-148) ------------------------------------------------
    @property
    def bandwidth(self) -> int:
        bw = self.inner_dim - 1
        if self._sub_diag is not None:
            bw += self.inner_dim
        return bw

    @property
    def batch_shape(self) -> torch.Size:
        return self._diag.shape[:-3]

    @property
    def inner_dim(self) -> int:
        return int(self._diag.shape[-2])

    @property
    def outer_dim(self) -> int:
        return int(self._diag.shape[-3])

    @property
    @boundary
    def block_diagonal(self) -> torch.Tensor:
        return self._diag

    @property
    @boundary
    def block_sub_diagonal(self) -> Optional[torch.Tensor]:
        return self._sub_diag

    # -- flattened, contiguous views handed to the C ABI ---------------------------------------
    def _flat(self):
        require_cuda(self._diag, "block diagonal")
        t, d = self.outer_dim, self.inner_dim
        b = _prod(self.batch_shape)
        diag = self._diag.reshape(b, t, d, d).contiguous()
        sub = None
        if self._sub_diag is not None:
            sub = self._sub_diag.reshape(b, t - 1, d, d).contiguous()
        return diag, sub, b, t, d

    # -- band view: API compatibility / debugging only (reference :90-98, :206-237) ------------
    @property
    @boundary
    def as_band(self) -> torch.Tensor:
        """Lower band ``[..., bandwidth+1, outer*inner]`` with ``band[r, j] = M[j+r, j]``."""
        t, d = self.outer_dim, self.inner_dim
        rows = self.bandwidth + 1
        dev = self._diag.device
        j = torch.arange(t * d, device=dev)
        r = torch.arange(rows, device=dev)
        k, c = j // d, j % d
        cr = c[None, :] + r[:, None]  # [rows, N]
        in_diag = cr < d
        band = torch.zeros(tuple(self.batch_shape) + (rows, t * d), dtype=self._diag.dtype, device=dev)
        kk = k[None, :].expand(rows, -1)
        cc = c[None, :].expand(rows, -1)
        vals = self._diag[..., kk, torch.clamp(cr, max=d - 1), cc]
        band = torch.where(in_diag, vals, band)
        if self._sub_diag is not None:
            in_sub = (~in_diag) & (cr < 2 * d) & (kk < t - 1)
            vals = self._sub_diag[..., torch.clamp(kk, max=t - 2), torch.clamp(cr - d, 0, d - 1), cc]
            band = torch.where(in_sub, vals, band)
        return band

    @boundary
    def to_dense(self) -> torch.Tensor:
        """Dense ``[..., outer*inner, outer*inner]`` (debugging; reference :150-173)."""
        t, d = self.outer_dim, self.inner_dim
        dense = torch.zeros(
            tuple(self.batch_shape) + (t * d, t * d), dtype=self._diag.dtype, device=self._diag.device
        )
        low = torch.tril(self._diag)
        for k in range(t):
            dense[..., k * d:(k + 1) * d, k * d:(k + 1) * d] = low[..., k, :, :]
            if self._sub_diag is not None and k + 1 < t:
                dense[..., (k + 1) * d:(k + 2) * d, k * d:(k + 1) * d] = self._sub_diag[..., k, :, :]
        if self._symmetric:
            dense = dense + torch.tril(dense, -1).transpose(-1, -2)
        return dense

    # -- right-hand-side broadcasting (reference :239-287) -------------------------------------
    def _prepare_right(self, right, skip_diag: bool = False):
        right = as_torch(right, self._diag.device)
        require_cuda(right, "right")
        t, d = self.outer_dim, self.inner_dim
        if right.dim() < 2 or tuple(right.shape[-2:]) != (t, d):
            raise ValueError(
                f"right must have shape [..., {t}, {d}], got {tuple(right.shape)}"
            )
        if right.dtype != self._diag.dtype:
            raise ValueError("right must have the dtype of the matrix")
        mb = tuple(self.batch_shape)
        rb = tuple(right.shape[:-2])
        try:
            fb = torch.broadcast_shapes(rb, mb)
        except RuntimeError as e:
            raise ValueError(f"right batch shape {rb} incompatible with matrix batch {mb}") from e
        diag, sub = self._diag, self._sub_diag
        mat_b = fb[len(fb) - len(mb):] if mb else ()
        if tuple(mat_b) != mb:  # the matrix itself has size-1 dims to expand (rare)
            diag = diag.expand(tuple(mat_b) + diag.shape[-3:])
            if sub is not None:
                sub = sub.expand(tuple(mat_b) + sub.shape[-3:])
        bm = _prod(mat_b)
        diag = None if skip_diag else diag.reshape(bm, t, d, d).contiguous()
        if sub is not None:
            sub = sub.reshape(bm, t - 1, d, d).contiguous()
        right = right.expand(tuple(fb) + (t, d)).reshape(_prod(fb), t, d).contiguous()
        return diag, sub, right, tuple(fb), bm

    @boundary
    def dense_mult(self, right, transpose_left: bool = False) -> torch.Tensor:
        """``L x``, ``Lᵀ x`` or (symmetric) ``M x`` (reference :175-199)."""
        diag, sub, rhs, fb, bm = self._prepare_right(right)
        t, d = self.outer_dim, self.inner_dim
        if needs_grad(diag, sub, rhs):
            return self._dense_mult_torch(diag, sub, rhs, bm, transpose_left).reshape(fb + (t, d))
        out = torch.empty_like(rhs)
        check(
            _lib.lib().mf_btd_dense_mult(
                dtype_code(diag.dtype), ptr(diag), ptr(sub), ptr(rhs), ptr(out), i64(rhs.shape[0]),
                i64(bm), i64(t), i64(d), int(bool(transpose_left)), int(bool(self._symmetric)),
                current_stream(),
            ),
            "mf_btd_dense_mult",
        )
        return out.reshape(fb + (t, d))

    def _dense_mult_torch(self, diag, sub, rhs, bm: int, transpose_left: bool) -> torch.Tensor:
        """The product in differentiable torch ops (a per-step map: no recursion), used when an operand
        requires a gradient."""
        n, t, d = rhs.shape
        x = rhs.reshape(n // bm, bm, t, d, 1)
        low = torch.tril(diag)
        if self._symmetric:
            low = low + torch.tril(diag, -1).transpose(-1, -2)
        elif transpose_left:
            low = low.transpose(-1, -2)
        y = low @ x
        if sub is not None:
            zero = torch.zeros_like(y[:, :, :1])
            down = torch.cat([zero, sub @ x[:, :, :-1]], dim=2)              # (L x)_k += sub_{k-1} x_{k-1}
            up = torch.cat([sub.transpose(-1, -2) @ x[:, :, 1:], zero], dim=2)  # (L^T x)_k += sub_k^T x_{k+1}
            if self._symmetric:
                y = y + down + up
            else:
                y = y + (up if transpose_left else down)
        return y[..., 0].reshape(n, t, d)

    @abc.abstractmethod
    def __add__(self, other):
        raise NotImplementedError

    def _added_blocks(self, other):
        if self._sub_diag is not None:
            sub = self._sub_diag
            if other.block_sub_diagonal is not None:
                sub = sub + other.block_sub_diagonal
        else:
            sub = other.block_sub_diagonal
        return self._diag + other.block_diagonal, sub


class LowerTriangularBlockTriDiagonal(BlockTriDiagonal):
    """Lower-triangular block-bidiagonal matrix (reference ``block_tri_diag.py:291-380``)."""

    def __init__(self, diagonal, sub_diagonal=None, unit_diagonal: bool = False) -> None:
        super().__init__(diagonal, symmetric=False, sub_diagonal=sub_diagonal)
        # unit_diagonal: the diagonal blocks are identities (``a_inv_block``); ``solve`` then skips
        # reading them (``diagonal`` may be an expanded, unmaterialised view)
        self._unit_diagonal = bool(unit_diagonal)

    @boundary
    def cholesky_of_block_inverses(self) -> torch.Tensor:
        """``chol((L_k L_kᵀ)⁻¹)`` for every diagonal block ``L_k`` (``kalman_filter.py:170-174``)."""
        require_cuda(self._diag, "block diagonal")
        d = self.inner_dim
        flat = self._diag.reshape(-1, d, d).contiguous()
        out = torch.empty_like(flat)
        check(
            _lib.lib().mf_block_chol_of_inverse(
                dtype_code(flat.dtype), ptr(flat), ptr(out), i64(flat.shape[0]), i64(d),
                current_stream()),
            "mf_block_chol_of_inverse",
        )
        return out.reshape(self._diag.shape)

    @boundary
    def block_diagonal_of_inverse(self) -> torch.Tensor:
        """Block diagonal of ``(L Lᵀ)⁻¹`` (reference :318-337)."""
        return self._inverse_subset(False)[0]

    def _inverse_subset(self, want_sub: bool):
        diag, sub, b, t, d = self._flat()
        out_d = torch.empty_like(diag)
        out_s = torch.empty_like(sub) if (want_sub and sub is not None) else None
        check(
            _lib.lib().mf_btd_inverse_subset(
                dtype_code(diag.dtype), ptr(diag), ptr(sub), ptr(out_d), ptr(out_s), i64(b), i64(t),
                i64(d), current_stream(),
            ),
            "mf_btd_inverse_subset",
        )
        bs = tuple(self.batch_shape)
        return out_d.reshape(bs + (t, d, d)), (None if out_s is None else out_s.reshape(bs + (t - 1, d, d)))

    @boundary
    def solve(self, right, transpose_left: bool = False) -> torch.Tensor:
        """``L⁻¹ x`` or ``L⁻ᵀ x`` (reference :339-351)."""
        diag, sub, rhs, fb, bm = self._prepare_right(right, skip_diag=self._unit_diagonal)
        t, d = self.outer_dim, self.inner_dim
        if needs_grad(diag, sub, rhs):
            return SolveFn.apply(diag, sub, rhs, bool(transpose_left)).reshape(fb + (t, d))
        out = torch.empty_like(rhs)
        check(
            _lib.lib().mf_btd_solve(
                dtype_code(rhs.dtype), ptr(diag), ptr(sub), ptr(rhs), ptr(out), i64(rhs.shape[0]),
                i64(bm), i64(t), i64(d), int(bool(transpose_left)), current_stream(),
            ),
            "mf_btd_solve",
        )
        return out.reshape(fb + (t, d))

    @boundary
    def abs_log_det(self) -> torch.Tensor:
        """``Σ log|L_nn|`` with shape ``batch_shape`` (reference :353-366)."""
        diag, _, b, t, d = self._flat()
        if needs_grad(diag):  # elementwise + reduction: differentiable torch ops
            ld = torch.log(torch.diagonal(diag, dim1=-2, dim2=-1).abs()).sum((-1, -2))
            return ld.reshape(tuple(self.batch_shape))
        out = torch.empty(b, dtype=diag.dtype, device=diag.device)
        check(
            _lib.lib().mf_btd_abs_log_det(
                dtype_code(diag.dtype), ptr(diag), ptr(out), i64(b), i64(t), i64(d), current_stream()
            ),
            "mf_btd_abs_log_det",
        )
        return out.reshape(tuple(self.batch_shape))

    @boundary
    def __add__(self, other: "LowerTriangularBlockTriDiagonal") -> "LowerTriangularBlockTriDiagonal":
        return LowerTriangularBlockTriDiagonal(*self._added_blocks(other))


class SymmetricBlockTriDiagonal(BlockTriDiagonal):
    """Symmetric block-tridiagonal matrix (reference ``block_tri_diag.py:384-545``)."""

    def __init__(self, diagonal, sub_diagonal=None) -> None:
        super().__init__(diagonal, symmetric=True, sub_diagonal=sub_diagonal)

    @boundary
    def __add__(self, other: "SymmetricBlockTriDiagonal") -> "SymmetricBlockTriDiagonal":
        return SymmetricBlockTriDiagonal(*self._added_blocks(other))

    @property
    @boundary
    def cholesky(self) -> LowerTriangularBlockTriDiagonal:
        """Block Cholesky ``L Lᵀ = M``; reads the lower triangle only (reference :423-436)."""
        return self.cholesky_and_solve(None)[0]

    @boundary
    def cholesky_and_solve(self, right=None, want_log_det: bool = False):
        """Fused sweep: factor ``M = L Lᵀ`` and, in the same pass, ``L⁻¹ right`` and ``log|L|``.

        One kernel replaces the reference sequence ``.cholesky`` -> ``.solve`` -> ``.abs_log_det``
        (``kalman_filter.py:220,244,251``).  ``right`` must have the matrix's batch shape.
        Returns ``(L, x_or_None, log_det_or_None)``.
        """
        diag, sub, b, t, d = self._flat()
        rhs = None
        if right is not None:
            rhs = as_torch(right, diag.device)
            if tuple(rhs.shape) != tuple(self.batch_shape) + (t, d) or rhs.dtype != diag.dtype:
                raise ValueError("right must have shape batch_shape + [outer_dim, inner_dim]")
            rhs = rhs.reshape(b, t, d).contiguous()
        if needs_grad(diag, sub, rhs):
            return self._cholesky_and_solve_diff(diag, sub, rhs, want_log_det)
        out_d = torch.empty_like(diag)
        out_s = torch.empty_like(sub) if sub is not None else None
        out_x = torch.empty_like(rhs) if rhs is not None else None
        logdet = torch.empty(b, dtype=diag.dtype, device=diag.device) if want_log_det else None
        info = torch.empty(b, dtype=torch.int32, device=diag.device)
        check(
            _lib.lib().mf_btd_cholesky(
                dtype_code(diag.dtype), ptr(diag), ptr(sub), ptr(rhs), ptr(out_d), ptr(out_s),
                ptr(out_x), ptr(logdet), ptr(info), i64(b), i64(t), i64(d), current_stream(),
            ),
            "mf_btd_cholesky",
        )
        _raise_if_failed(info, "SymmetricBlockTriDiagonal.cholesky")
        bs = tuple(self.batch_shape)
        chol = LowerTriangularBlockTriDiagonal(
            out_d.reshape(bs + (t, d, d)), None if out_s is None else out_s.reshape(bs + (t - 1, d, d))
        )
        x = None if out_x is None else out_x.reshape(bs + (t, d))
        ld = None if logdet is None else logdet.reshape(bs)
        return chol, x, ld

    @boundary
    def _cholesky_and_solve_diff(self, diag, sub, rhs, want_log_det: bool):
        """The same three results from the differentiable primitives (adjoint sweeps, autograd.py)."""
        t, d = self.outer_dim, self.inner_dim
        ld, ls, info = CholeskyFn.apply(diag, sub)
        _raise_if_failed(info, "SymmetricBlockTriDiagonal.cholesky")
        bs = tuple(self.batch_shape)
        chol = LowerTriangularBlockTriDiagonal(
            ld.reshape(bs + (t, d, d)), None if ls is None else ls.reshape(bs + (t - 1, d, d)))
        x = None if rhs is None else SolveFn.apply(ld, ls, rhs, False).reshape(bs + (t, d))
        logdet = chol.abs_log_det() if want_log_det else None
        return chol, x, logdet

    def upper_diagonal_lower(
        self,
    ) -> Tuple[LowerTriangularBlockTriDiagonal, LowerTriangularBlockTriDiagonal]:
        """``UDUᵀ`` factorisation: returns ``(Uᵀ, chol_D)`` (reference :438-545)."""
        assert self._sub_diag is not None
        diag, sub, b, t, d = self._flat()
        out_u = torch.empty_like(sub)
        out_cd = torch.empty_like(diag)
        info = torch.empty(b, dtype=torch.int32, device=diag.device)
        check(
            _lib.lib().mf_btd_upper_diagonal_lower(
                dtype_code(diag.dtype), ptr(diag), ptr(sub), ptr(out_u), ptr(out_cd), ptr(info),
                i64(b), i64(t), i64(d), current_stream(),
            ),
            "mf_btd_upper_diagonal_lower",
        )
        _raise_if_failed(info, "SymmetricBlockTriDiagonal.upper_diagonal_lower")
        bs = tuple(self.batch_shape)
        eye = torch.eye(d, dtype=diag.dtype, device=diag.device).expand(bs + (t, d, d))
        return (
            LowerTriangularBlockTriDiagonal(eye, out_u.reshape(bs + (t - 1, d, d)), unit_diagonal=True),
            LowerTriangularBlockTriDiagonal(out_cd.reshape(bs + (t, d, d))),
        )
