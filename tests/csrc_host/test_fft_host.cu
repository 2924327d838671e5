// Host-side check of the register DFT templates and the two-stage row FFT index algebra
// (compiled with nvcc for the host only; no GPU needed).  Exit code 0 = all good.
#include <cstdio>
#include <cmath>
#include <vector>
#include <complex>
#include "../../ptyrad_b200/csrc/rowfft.cuh"

using namespace ptyb;
typedef std::complex<double> cdbl;

template <int R, int DIR> double check_dft() {
    float2 v[R];
    cdbl x[R];
    for (int n = 0; n < R; ++n) { x[n] = cdbl(std::sin(1.3 * n + 0.2), std::cos(0.7 * n * n)); v[n] = make_float2((float)x[n].real(), (float)x[n].imag()); }
    Dft<R, DIR>::run(v);
    double err = 0, nrm = 0;
    for (int k = 0; k < R; ++k) {
        cdbl s = 0;
        for (int n = 0; n < R; ++n) s += x[n] * std::polar(1.0, DIR * 2 * M_PI * n * k / R);
        err += std::norm(s - cdbl(v[k].x, v[k].y)); nrm += std::norm(s);
    }
    return std::sqrt(err / nrm);
}

template <int N1, int N2> double check_row() {
    typedef RowFFT<N1, N2> F;
    const int N = F::N;
    std::vector<float2> tw(N), row(F::RS), row0(F::RS);
    for (int n = 0; n < N; ++n) tw[n] = make_float2((float)std::cos(-2 * M_PI * n / N), (float)std::sin(-2 * M_PI * n / N));
    std::vector<cdbl> x(N);
    for (int n = 0; n < N; ++n) { x[n] = cdbl(std::sin(0.37 * n) + 0.1 * n / N, std::cos(1.1 * n + 0.3)); row[F::addr(n)] = make_float2((float)x[n].real(), (float)x[n].imag()); }
    row0 = row;
    for (int j = 0; j < N2; ++j) F::fwd_stage1(row.data(), j, tw.data());
    for (int k1 = 0; k1 < N1; ++k1) F::fwd_stage2(row.data(), k1);
    double err = 0, nrm = 0;
    for (int q = 0; q < N; ++q) {
        cdbl s = 0;
        for (int n = 0; n < N; ++n) s += x[n] * std::polar(1.0, -2 * M_PI * ((long long)n * q % N) / N);
        float2 g = row[F::apos(q)];
        err += std::norm(s - cdbl(g.x, g.y)); nrm += std::norm(s);
    }
    double e1 = std::sqrt(err / nrm);
    for (int k1 = 0; k1 < N1; ++k1) F::inv_stage2(row.data(), k1, tw.data());
    for (int j = 0; j < N2; ++j) F::inv_stage1(row.data(), j);
    err = 0; nrm = 0;
    for (int n = 0; n < N; ++n) {
        float2 g = row[F::addr(n)];
        err += std::norm(x[n] * double(N) - cdbl(g.x, g.y)); nrm += std::norm(x[n] * double(N));
    }
    double e2 = std::sqrt(err / nrm);
    return e1 > e2 ? e1 : e2;
}

// the register-fed / register-draining outer stages used by the general kernels (stages A and C) against the slab versions
template <int N1, int N2> double check_row_regs() {
    typedef RowFFT<N1, N2> F;
    const int N = F::N;
    std::vector<float2> tw(N), row(F::RS), ref(F::RS);
    for (int n = 0; n < N; ++n) tw[n] = make_float2((float)std::cos(-2 * M_PI * n / N), (float)std::sin(-2 * M_PI * n / N));
    std::vector<float2> x(N);
    for (int n = 0; n < N; ++n) { x[n] = make_float2((float)(std::sin(0.37 * n) + 0.1 * n / N), (float)std::cos(1.1 * n + 0.3)); ref[F::addr(n)] = x[n]; }
    // forward: A from registers, C into registers; reference = slab stages
    for (int j = 0; j < N2; ++j) {
        float2 v[N1];
        for (int k = 0; k < N1; ++k) v[k] = x[j + N2 * k];
        F::fwd_stage1_regs(row.data(), j, v, tw.data());
        F::fwd_stage1(ref.data(), j, tw.data());
    }
    std::vector<float2> X(N);
    for (int k1 = 0; k1 < N1; ++k1) {
        float2 v[N2];
        F::fwd_stage2_regs(row.data(), k1, v);
        for (int k2 = 0; k2 < N2; ++k2) X[k1 + N1 * k2] = v[k2];
        F::fwd_stage2(ref.data(), k1);
    }
    double err = 0, nrm = 0;
    for (int q = 0; q < N; ++q) { float2 g = ref[F::apos(q)]; err += std::norm(cdbl(g.x - X[q].x, g.y - X[q].y)); nrm += std::norm(cdbl(g.x, g.y)); }
    // inverse: A from registers (frequency order), C into registers (natural order)
    for (int k1 = 0; k1 < N1; ++k1) {
        float2 v[N2];
        for (int k2 = 0; k2 < N2; ++k2) v[k2] = X[k1 + N1 * k2];
        F::inv_stage2_regs(row.data(), k1, v, tw.data());
    }
    for (int j = 0; j < N2; ++j) {
        float2 v[N1];
        F::inv_stage1_regs(row.data(), j, v);
        for (int k = 0; k < N1; ++k) {
            const float2 w = x[j + N2 * k];
            err += std::norm(cdbl(v[k].x / N - w.x, v[k].y / N - w.y)); nrm += std::norm(cdbl(w.x, w.y));
        }
    }
    return std::sqrt(err / nrm);
}

int main() {
    int bad = 0;
#define CHK(expr, tol) { double e = (expr); printf("%-28s %.3e\n", #expr, e); if (!(e < tol)) { bad++; printf("   FAIL\n"); } }
    CHK((check_dft<2, -1>()), 1e-6); CHK((check_dft<3, -1>()), 1e-6); CHK((check_dft<4, -1>()), 1e-6);
    CHK((check_dft<6, -1>()), 1e-6); CHK((check_dft<8, -1>()), 1e-6); CHK((check_dft<12, -1>()), 1e-6);
    CHK((check_dft<16, -1>()), 1e-6); CHK((check_dft<32, -1>()), 1e-6); CHK((check_dft<24, 1>()), 1e-6);
    CHK((check_dft<3, 1>()), 1e-6); CHK((check_dft<4, 1>()), 1e-6); CHK((check_dft<8, 1>()), 1e-6);
    CHK((check_dft<16, 1>()), 1e-6); CHK((check_dft<32, 1>()), 1e-6); CHK((check_dft<12, 1>()), 1e-6);
    CHK((check_row<4, 4>()), 1e-6); CHK((check_row<8, 4>()), 1e-6); CHK((check_row<8, 6>()), 1e-6);
    CHK((check_row<8, 8>()), 1e-6); CHK((check_row<12, 8>()), 1e-6); CHK((check_row<16, 8>()), 1e-6);
    CHK((check_row<16, 12>()), 1e-6); CHK((check_row<16, 16>()), 1e-6);
    CHK((check_row_regs<4, 4>()), 1e-6); CHK((check_row_regs<8, 6>()), 1e-6); CHK((check_row_regs<12, 8>()), 1e-6);
    CHK((check_row_regs<16, 8>()), 1e-6); CHK((check_row_regs<16, 12>()), 1e-6); CHK((check_row_regs<16, 16>()), 1e-6);
    printf(bad ? "FAILED %d\n" : "ALL OK\n", bad);
    return bad;
}
