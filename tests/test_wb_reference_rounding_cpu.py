"""CPU side of tests/test_gpu_wb_reference_rounding.py: the numpy fp32 restatement of the reference graph's PER-PIXEL
write-back gradient arithmetic that the GPU tests compare the kernel with is itself pinned here -- to oracle/st_oracle.c,
which is bit-exact with the gradient subgraph of the reference's saved graph (tests/test_reference_graph.py)."""
import numpy as np

from oracle import c_oracle as C
from tests import test_gpu_wb_reference_rounding as G

F = np.float32


def test_per_pixel_restatement_equals_the_graph_exact_c_oracle():
    B = 48
    win, thi, z, stop, d = G._spiky_case(B, 11)
    dth, scale_th, dzs, scale_z, sample, terms = G._graph_per_pixel(win, thi, z, d)
    # forward samples (residues included): bit for bit
    plain = C.st_forward(win.reshape(B, 28, 28, 1), thi.reshape(B, 2, 3), (50, 50)).reshape(B, 50, 50)
    assert np.array_equal(sample, plain)
    # d theta_inv: the C oracle sums the same per-pixel values in pixel order in fp32; the restatement sums them in fp64
    up = (d * z[:, None, None]).astype(F)
    _, cth = C.st_backward(win.reshape(B, 28, 28, 1), thi.reshape(B, 2, 3), up.reshape(B, 50, 50, 1), need_dU=False)
    cth = cth.reshape(B, 6).astype(np.float64)
    assert (np.abs(cth - dth) <= 2e-5 * scale_th + 1e-6).all(), np.abs(cth - dth).max()
    assert np.abs(d).max() > 1e6 and scale_th.max() > 100.0    # the fixture has the 1e9 / 64 spikes and un-cancelled terms


def test_rows_clipped_in_y_contribute_exactly_zero_in_the_graphs_order():
    """What st_wb_bwd_ref and the forward kernels rely on: a canvas pixel whose ROW is clipped has sample == +0.0 and
    per-pixel dx == dy == 0.0 exactly in the order of the reference's add_n / AddN (adjacent terms cancel), whereas a
    pixel whose COLUMN alone is clipped leaves the residues."""
    B = 48
    win, thi, z, stop, d = G._spiky_case(B, 5)
    xt, j0, j1, cw1, cw0 = G._tables(thi[:, 0], thi[:, 2], 50, 28)
    yt, i0, i1, rw1, rw0 = G._tables(thi[:, 4], thi[:, 5], 50, 28)
    row_clipped = (i0 == i1)[:, :, None] & np.ones((1, 1, 50), bool)
    col_only = (j0 == j1)[:, None, :] & ~row_clipped
    plain = C.st_forward(win.reshape(B, 28, 28, 1), thi.reshape(B, 2, 3), (50, 50)).reshape(B, 50, 50)
    assert row_clipped.mean() > 0.2 and col_only.mean() > 0.1
    assert not plain[row_clipped].any() and not np.signbit(plain[row_clipped]).any()
    assert (plain[col_only] != 0).mean() > 0.3
    # per-pixel dx, dy of the row-clipped pixels: give ONLY them an upstream gradient and compare with zero
    up = np.where(row_clipped, F(1e9), F(0)).astype(F)
    _, cth = C.st_backward(win.reshape(B, 28, 28, 1), thi.reshape(B, 2, 3), up.reshape(B, 50, 50, 1), need_dU=False)
    assert not cth.any()
    up = np.where(col_only, F(1e9), F(0)).astype(F)
    _, cth = C.st_backward(win.reshape(B, 28, 28, 1), thi.reshape(B, 2, 3), up.reshape(B, 50, 50, 1), need_dU=False)
    assert np.abs(cth).max() > 1.0
