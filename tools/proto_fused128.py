"""NumPy model of the fused-128 register/shared-memory FFT data flow (index algebra check, no GPU).

512 threads x 32 complex registers hold one 128x128 tile.
  layout R: thread t -> x = t & 127, yl = t >> 7 ; v[k] = psi[yl + 4k][x]
  layout F: thread t -> w2 = t >> 5, l = t & 31, rsel = l >> 4, q = (l >> 2) & 3, vv = l & 3 ;
            v[u] = X[ky = w2 + 16 rsel + 32 q][kx = vv + 4u]
forward (R -> F):  DFT32 over k | E1 exchange (CTA) | y-twiddle, 4x4 DFT, x-twiddle | E2 exchange (warp) | DFT32 over j
inverse (F -> R):  the same stages reversed with conjugate twiddles (unnormalised: x N^2).
"""
import numpy as np

NT, NR = 512, 32
t = np.arange(NT)
W = lambda n, e: np.exp(-2j * np.pi * e / n)


def e2_pos(e16, j):
    return e16 * 32 + (j ^ e16)


def fwd(v):                      # v: (512,32) layout R
    # stage 1: DFT32 over registers
    k = np.arange(32)
    D32 = W(32, np.outer(k, k))
    v = v @ D32                  # v[t, r]
    x, yl = t & 127, t >> 7
    E = np.zeros(32 * 4 * 128, complex)
    for r in range(32):
        E[(r * 4 + yl) * 128 + x] = v[:, r]
    # stage 2 (T2): w2 = t>>5, j = t&31
    w2, j = t >> 5, t & 31
    out = np.zeros((NT, 2, 4, 4), complex)
    for rsel in range(2):
        r = w2 + 16 * rsel
        inn = np.zeros((NT, 4, 4), complex)
        for yl_ in range(4):
            for s in range(4):
                inn[:, yl_, s] = E[(r * 4 + yl_) * 128 + j + 32 * s] * W(128, yl_ * r)
        D4 = W(4, np.outer(np.arange(4), np.arange(4)))
        o = np.einsum('tys,yq,sv->tqv', inn, D4, D4)
        for vv in range(4):
            o[:, :, vv] *= W(128, j * vv)[:, None]
        out[:, rsel] = o
    # E2: warp-local transpose inside the warp's own two 4 KB chunks of E
    E2 = np.zeros_like(E)
    for rsel in range(2):
        base = (w2 + 16 * rsel) * 512
        for q in range(4):
            for vv in range(4):
                e16 = q * 4 + vv
                E2[base + e2_pos(e16, j)] = out[:, rsel, q, vv]
    l = t & 31
    rsel, e16 = l >> 4, l & 15
    base = (w2 + 16 * rsel) * 512
    v3 = np.zeros((NT, 32), complex)
    for jj in range(32):
        v3[:, jj] = E2[base + e2_pos(e16, jj)]
    return v3 @ D32              # v[t, u]


def inv(v):                      # v: (512,32) layout F -> layout R, unnormalised
    k = np.arange(32)
    ID32 = np.conj(W(32, np.outer(k, k)))
    v = v @ ID32                 # over u -> j
    w2, l = t >> 5, t & 31
    rsel, e16 = l >> 4, l & 15
    E2 = np.zeros(32 * 4 * 128, complex)
    base = (w2 + 16 * rsel) * 512
    for jj in range(32):
        E2[base + e2_pos(e16, jj)] = v[:, jj]
    j = t & 31
    E = np.zeros_like(E2)
    ID4 = np.conj(W(4, np.outer(np.arange(4), np.arange(4))))
    for rs in range(2):
        r = w2 + 16 * rs
        b2 = r * 512
        inn = np.zeros((NT, 4, 4), complex)
        for q in range(4):
            for vv in range(4):
                inn[:, q, vv] = E2[b2 + e2_pos(q * 4 + vv, j)] * np.conj(W(128, j * vv))
        o = np.einsum('tqv,qy,vs->tys', inn, ID4, ID4)
        for yl_ in range(4):
            for s in range(4):
                E[(r * 4 + yl_) * 128 + j + 32 * s] = o[:, yl_, s] * np.conj(W(128, yl_ * r))
    x, yl = t & 127, t >> 7
    vr = np.zeros((NT, 32), complex)
    for r in range(32):
        vr[:, r] = E[(r * 4 + yl) * 128 + x]
    return vr @ ID32


def to_R(a):                     # (128,128) -> (512,32)
    x, yl = t & 127, t >> 7
    return np.stack([a[yl + 4 * k, x] for k in range(32)], 1)


def from_F(v):
    w2, l = t >> 5, t & 31
    rsel, q, vv = l >> 4, (l >> 2) & 3, l & 3
    ky = w2 + 16 * rsel + 32 * q
    out = np.zeros((128, 128), complex)
    for u in range(32):
        out[ky, vv + 4 * u] = v[:, u]
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    a = rng.standard_normal((128, 128)) + 1j * rng.standard_normal((128, 128))
    X = from_F(fwd(to_R(a)))
    print("fwd err", np.abs(X - np.fft.fft2(a)).max() / np.abs(X).max())
    back = inv(fwd(to_R(a)))
    print("roundtrip err", np.abs(back - to_R(a) * 128 * 128).max() / (128 * 128))
