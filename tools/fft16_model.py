#!/usr/bin/env python
"""Thread-level model of k_idct_r16_field (pic-trapped-plasma_b200/csrc/ptp_solve_wide.cu): the inverse DCT-I of one
4096-node grid row through three register-resident radix-16 stages. Every index expression, twiddle and shared-memory
address below is the one the CUDA kernel uses; the model checks (i) the result against the DCT-I sum, (ii) that no
quarter-warp of 16-byte shared-memory accesses hits a bank group twice. Run: python tools/fft16_model.py"""
import numpy as np

N = 4096
T = 256
RS = 257                      # row stride (16-byte elements) of the [16][RS] exchange buffer
tw = np.exp(-1j * np.pi * np.arange(N) / N)        # fftTw[j] = exp(-i pi j / N)


def tw2(j):                   # exp(-i pi j / N) for 0 <= j < 2N from the half-circle table
    return tw[j] if j < N else -tw[j - N]


C8, S8, R2 = np.cos(np.pi / 8), np.sin(np.pi / 8), np.sqrt(0.5)
W16 = {0: 1, 1: C8 - 1j * S8, 2: R2 - 1j * R2, 3: S8 - 1j * C8, 4: -1j, 6: -R2 - 1j * R2, 9: -(C8 - 1j * S8)}


def fft4(a0, a1, a2, a3):
    s0, s1, s2, s3 = a0 + a2, a0 - a2, a1 + a3, a1 - a3
    m = -1j * s3                  # (s3.y, -s3.x)
    return s0 + s2, s1 + m, s0 - s2, s1 - m


def fft16(x):
    """y[k] = sum_j x[j] exp(-2 pi i j k / 16), natural order in and out (4 x 4 decomposition j = 4 j1 + j0, k = k0 + 4 k1)."""
    t = [[None] * 4 for _ in range(4)]
    for j0 in range(4):
        t[j0] = list(fft4(x[j0], x[4 + j0], x[8 + j0], x[12 + j0]))     # over j1 -> k0
    for j0 in range(1, 4):
        for k0 in range(1, 4):
            t[j0][k0] = t[j0][k0] * W16[j0 * k0]
    y = [None] * 16
    for k0 in range(4):
        y[k0], y[k0 + 4], y[k0 + 8], y[k0 + 12] = fft4(t[0][k0], t[1][k0], t[2][k0], t[3][k0])   # over j0 -> k1
    return y


def powers(w1):
    """w[k] = w1^k, k = 0..15, by binary splitting (depth <= 4 products)."""
    w = [1, w1] + [None] * 14
    w[2] = w1 * w1
    w[3] = w[2] * w1
    w[4] = w[2] * w[2]
    for k in (5, 6, 7):
        w[k] = w[4] * w[k - 4]
    w[8] = w[4] * w[4]
    for k in range(9, 16):
        w[k] = w[8] * w[k - 8]
    return w


class Smem:
    def __init__(self, n):
        self.a = np.zeros(n, dtype=complex)
        self.worst = 1

    def access(self, addr_of_thread):
        """addr_of_thread: list of T element indices (one 16-byte access per thread). Checks quarter-warp bank groups."""
        for q in range(0, T, 8):
            groups = [addr_of_thread[q + l] % 8 for l in range(8)]
            self.worst = max(self.worst, max(groups.count(g) for g in set(groups)))


def idct_row(a):
    S = Smem(16 * RS)
    # ---- stage A: thread t = 16 n1 + n0 owns samples n = t + 256 j ----------------------------------------------
    regs = []
    for t in range(T):
        x = []
        for j in range(16):
            n = t + 256 * j
            i0, i1 = 2 * n, 2 * n + 1
            x.append(a[i0 if i0 <= N else 2 * N - i0] + 1j * a[i1 if i1 <= N else 2 * N - i1])
        y = fft16(x)
        w = powers(tw2(2 * t))                    # W_N^t
        regs.append([y[k] * w[k] for k in range(16)])
    for k0 in range(16):
        addr = [(t >> 4) * RS + k0 * 16 + (t & 15) for t in range(T)]
        S.access(addr)
        for t in range(T):
            S.a[addr[t]] = regs[t][k0]
    # ---- stage B: thread v = 16 k0 + n0 ------------------------------------------------------------------------
    regs = []
    for n1 in range(16):
        S.access([n1 * RS + v for v in range(T)])
    for v in range(T):
        x = [S.a[n1 * RS + v] for n1 in range(16)]
        y = fft16(x)
        n0 = v & 15
        w = powers(tw2(32 * n0))                  # W_256^n0
        regs.append([y[k] * w[k] for k in range(16)])
    # (barrier: every read above precedes every write below)
    for k1 in range(16):
        addr = [(v & 15) * RS + (v >> 4) + 16 * k1 for v in range(T)]
        S.access(addr)
        for v in range(T):
            S.a[addr[v]] = regs[v][k1]
    # ---- stage C: thread u = k0 + 16 k1 -------------------------------------------------------------------------
    regs = []
    for n0 in range(16):
        S.access([n0 * RS + u for u in range(T)])
    for u in range(T):
        regs.append(fft16([S.a[n0 * RS + u] for n0 in range(16)]))
    for k2 in range(16):
        addr = [u + 256 * k2 for u in range(T)]
        S.access(addr)
        for u in range(T):
            S.a[addr[u]] = regs[u][k2]               # Z[k0 + 16 k1 + 256 k2] in natural order
    # ---- read-out: thread t owns k = t + 256 j (j < 8) and N - k; thread 0 also k = N / 2 ---------------------------
    out = np.zeros(N + 1)
    a0, aN = a[0], a[N]

    def emit(k, A, B, w):
        dx, dy = A.real - B.real, A.imag + B.imag
        X = 0.5 * (A.real + B.real) + 0.5 * (dy * w.real + dx * w.imag)
        out[k] = 0.5 * X + 0.5 * (a0 + (-aN if k & 1 else aN))

    for j in range(8):
        S.access([t + 256 * j for t in range(T)])
        S.access([(N - (t + 256 * j)) & (N - 1) for t in range(T)])
        for t in range(T):
            k = t + 256 * j
            A, B = S.a[k], S.a[(N - k) & (N - 1)]
            emit(k, A, B, tw[k])
            emit(N - k, B, A, -np.conj(tw[k]))       # W_2N^{N-k} = -conj W_2N^k
    A = S.a[N // 2]
    emit(N // 2, A, A, tw[N // 2])
    return out, S.worst


def main():
    rng = np.random.default_rng(1)
    a = rng.standard_normal(N + 1)
    from scipy.fft import dct
    k = np.arange(N + 1)
    ref = 0.5 * (dct(a, type=1) + a[0] + np.where(k & 1, -a[N], a[N]))     # sum_m a_m cos(pi m k / N)
    x = [complex(v, w) for v, w in rng.standard_normal((16, 2))]
    assert np.allclose(fft16(x), np.fft.fft(np.array(x)), rtol=0, atol=1e-13)
    out, worst = idct_row(a)
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    print("rel-L2 vs DCT-I sum: %.2e   worst quarter-warp bank-group multiplicity: %d" % (err, worst))
    assert err < 1e-14 and worst == 1


if __name__ == "__main__":
    main()
