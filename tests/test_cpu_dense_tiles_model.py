"""Host model of the register-tiled L D L^T of the reduced system (vc_imu_mega.cuh: dense_ldlt_tiles).  The device loop
has no bounds checks: the matrix is zero-padded to 16 T rows, every thread (ti, tk) updates ALL its elements
(16 ii + ti, 16 kk + tk), kk <= ii, with column j at every step, and only finished columns go back to shared memory.
The claim that makes this legal — an update that "should not happen" (row or column already finished, padding, above
the diagonal) only ever lands in an element nobody reads again — is checked here by running exactly that schedule in
numpy and comparing what went back to shared memory with a plain L D L^T."""
import numpy as np
import pytest


def _tiles(N):
    t = (N + 1 + 15) // 16
    return 6 if t <= 6 else 7 if t <= 7 else 9 if t <= 9 else 0


def _device_schedule(H, g):
    """H: N x N SPD, g: right-hand side.  Returns (S, wd): S[i][j] = u_ij for i >= j (row N = right-hand side row),
    wd[j] = 1 / d_j, as the device leaves them in shared memory."""
    N = H.shape[0]
    T = _tiles(N)
    RT, LD = 16 * T, 16 * T + 1
    S = np.zeros((RT, LD))
    S[:N, :N] = H          # both triangles, as the assembly writes them
    S[N, :N] = g           # the right-hand side rides along as row N
    wd = np.zeros(N)
    # every thread's registers: R[ti, tk, ii, kk] = S[16 ii + ti, 16 kk + tk] for kk <= ii (other tiles unused)
    ii, kk = np.meshgrid(np.arange(T), np.arange(T), indexing="ij")
    R = np.zeros((16, 16, T, T))
    for ti in range(16):
        for tk in range(16):
            R[ti, tk] = np.where(kk <= ii, S[16 * ii + ti, 16 * kk + tk], np.nan)  # nan: a register that does not exist
    wd[0] = 1.0 / S[0, 0]
    for j in range(N):
        jt = j // 16
        wj = wd[j]
        for ti in range(16):
            ci = S[16 * np.arange(T) + ti, j] * wj          # no bounds check, no i > j check
            for tk in range(16):
                ck = S[16 * np.arange(T) + tk, j]           # no k > j check
                upd = np.outer(ci, ck)
                live = (kk <= ii) & (ii >= jt) & (kk >= jt)  # the tiles the unrolled loops touch
                R[ti, tk][live] -= upd[live]
        jn = j + 1
        if jn < N:  # the owners of column j + 1 put it back: ALL rows of the tiles from kk_n on, dead ones included
            kkn, tk = jn // 16, jn % 16
            for ti in range(16):
                for i2 in range(kkn, T):
                    S[16 * i2 + ti, jn] = R[ti, tk, i2, kkn]
            wd[jn] = 1.0 / R[jn % 16, tk, kkn, kkn]
    return S, wd


@pytest.mark.parametrize("N", [63, 81, 90, 99, 135])
def test_padded_schedule_equals_ldlt(N):
    rng = np.random.default_rng(N)
    A = rng.standard_normal((N + 20, N))
    H = A.T @ A + 0.1 * np.eye(N)
    g = rng.standard_normal(N)
    S, wd = _device_schedule(H, g)
    # plain right-looking L D L^T with the right-hand side as an extra row
    M = np.vstack([H, g[None, :]])
    d = np.zeros(N)
    for j in range(N):
        d[j] = M[j, j]
        col = M[j + 1:, j].copy()
        M[j + 1:, j + 1:] -= np.outer(col / d[j], M[j + 1:N, j])
    for j in range(N):
        assert np.allclose(S[j:N + 1, j], M[j:, j], rtol=1e-12, atol=1e-12 * abs(H).max()), j
    assert np.allclose(wd, 1.0 / d, rtol=1e-12)
    # and the back-substitution the device runs on it solves H x = g
    x = S[N, :N].copy()
    for i in range(N - 1, -1, -1):
        x[i] *= wd[i]
        x[:i] -= S[i, :i] * x[i]
    assert np.allclose(H @ x, g, rtol=1e-8, atol=1e-8)
