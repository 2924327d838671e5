"""NumPy model of the index algebra of csrc/fused256.cuh (the 8-CTA-cluster 256x256 2-D FFT), checked against np.fft.fft2.

A 256x256 complex wave is shared by the 8 CTAs of a cluster, 256 threads each, 32 complex values per thread:

  layout R (real space)  CTA c, thread t: x = 32c + (t & 31), yl = t >> 5;      v[k] = psi[yl + 8k][x]
  layout M (middle)      CTA c, thread t: a = t >> 3, xl = t & 7;               v[4b + i] = Y[32c + xl + 8i][ky = a + 32b]
  layout F1              CTA c', thread t: a = t >> 3, xl = t & 7;              v[k] = Y[xl + 8k][ky = a + 32c']
  layout F (Fourier)     CTA c', thread t: ky = 32c' + (t >> 3), lx = t & 7;    v[u] = X[ky][kx = lx + 8u]

  forward: DFT32 over k (registers) -> exchange Ea (CTA wide, shared memory) -> twiddle W256^(yl a), DFT8 over yl (layout M)
           -> transposition T between the CTAs of the cluster (16-byte vectors = register pairs; through L2 or DSMEM) -> layout F1
           -> DFT32 over k (registers) -> exchange Eb (warp local) -> twiddle W256^(xl a2), DFT8 over xl -> layout F
  inverse: the same stages backwards with conjugate twiddles (T is its own mirror: the same store/load index functions).

Run:  python tools/proto_fused256.py
"""
import numpy as np

N, C, T, R = 256, 8, 256, 32
W = lambda n, s: np.exp(s * 2j * np.pi * n / 256.0)


def to_R(psi):
    v = np.zeros((C, T, R), complex)
    for c in range(C):
        for t in range(T):
            x, yl = 32 * c + (t & 31), t >> 5
            v[c, t, :] = psi[yl + 8 * np.arange(R), x]
    return v


def from_R(v):
    psi = np.zeros((N, N), complex)
    for c in range(C):
        for t in range(T):
            x, yl = 32 * c + (t & 31), t >> 5
            psi[yl + 8 * np.arange(R), x] = v[c, t, :]
    return psi


def from_F(v):
    X = np.zeros((N, N), complex)
    for c in range(C):
        for t in range(T):
            ky, lx = 32 * c + (t >> 3), t & 7
            X[ky, lx + 8 * np.arange(R)] = v[c, t, :]
    return X


def to_F(X):
    v = np.zeros((C, T, R), complex)
    for c in range(C):
        for t in range(T):
            ky, lx = 32 * c + (t >> 3), t & 7
            v[c, t, :] = X[ky, lx + 8 * np.arange(R)]
    return v


def dft(v, s):          # unnormalised DFT over the last axis, sign s
    return np.fft.fft(v, axis=-1) if s < 0 else np.fft.ifft(v, axis=-1) * v.shape[-1]


def transpose_T(v):
    """register pairs q = (2q, 2q+1): CTA c stores pair q at slot [dest = q >> 1][2c + (q & 1)][t]; CTA c loads pair q from [c][q][t]"""
    scratch = np.zeros((C, 16, T, 2), complex)
    for c in range(C):
        for q in range(16):
            scratch[q >> 1, 2 * c + (q & 1), :, :] = v[c, :, 2 * q:2 * q + 2]
    out = np.zeros_like(v)
    for c in range(C):
        for q in range(16):
            out[c, :, 2 * q:2 * q + 2] = scratch[c, q, :, :]
    return out


def fft2_R_to_F(v, s=-1):
    v = dft(v, s)                                            # DFT32 over k -> a
    m = np.zeros_like(v)
    for c in range(C):
        E = np.zeros((32, 8, 32), complex)                   # Ea[a][yl][xloc]
        for t in range(T):
            E[:, t >> 5, t & 31] = v[c, t, :]
        for t in range(T):
            a, xl = t >> 3, t & 7
            for i in range(4):
                col = E[a, :, xl + 8 * i] * W(np.arange(8) * a, s)
                m[c, t, 4 * np.arange(8) + i] = dft(col, s)  # -> b
    v = transpose_T(m)                                       # layout F1
    v = dft(v, s)                                            # DFT32 over k -> a2
    f = np.zeros_like(v)
    for c in range(C):
        for w in range(8):
            E = np.zeros((32, 4, 8), complex)                # Eb[a2][a_l][xl]  (warp local)
            for l in range(32):
                E[:, l >> 3, l & 7] = v[c, 32 * w + l, :]
            for l in range(32):
                al, lx = l >> 3, l & 7
                for i in range(4):
                    a2 = lx + 8 * i
                    col = E[a2, al, :] * W(np.arange(8) * a2, s)
                    f[c, 32 * w + l, i + 4 * np.arange(8)] = dft(col, s)
    return f


def fft2_F_to_R(f, s=+1):
    v = np.zeros_like(f)
    for c in range(C):
        for w in range(8):
            E = np.zeros((32, 4, 8), complex)
            for l in range(32):
                al, lx = l >> 3, l & 7
                for i in range(4):
                    a2 = lx + 8 * i
                    E[a2, al, :] = dft(f[c, 32 * w + l, i + 4 * np.arange(8)], s) * W(np.arange(8) * a2, s)
            for l in range(32):
                v[c, 32 * w + l, :] = E[:, l >> 3, l & 7]
    v = dft(v, s)                                            # over a2 -> k
    m = transpose_T(v)                                       # layout M
    v = np.zeros_like(m)
    for c in range(C):
        E = np.zeros((32, 8, 32), complex)
        for t in range(T):
            a, xl = t >> 3, t & 7
            for i in range(4):
                E[a, :, xl + 8 * i] = dft(m[c, t, 4 * np.arange(8) + i], s) * W(np.arange(8) * a, s)
        for t in range(T):
            v[c, t, :] = E[:, t >> 5, t & 31]
    return dft(v, s)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    psi = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    F = fft2_R_to_F(to_R(psi))
    ref = np.fft.fft2(psi)
    print("forward  R->F vs np.fft.fft2 :", np.abs(from_F(F) - ref).max() / np.abs(ref).max())
    back = fft2_F_to_R(F)
    print("inverse  F->R round trip     :", np.abs(from_R(back) / N**2 - psi).max())
    X = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    inv = from_R(fft2_F_to_R(to_F(X))) / N**2
    print("inverse  F->R vs np.fft.ifft2:", np.abs(inv - np.fft.ifft2(X)).max() / np.abs(inv).max())
