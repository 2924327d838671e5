// Register-resident small DFTs with compile-time twiddles (sm_100a; also compiles for the host so the
// index algebra can be unit-tested without a GPU).
//
// Dft<R, DIR>::run(v) transforms v[0..R) in place, natural order in and out,
//   X[k] = sum_n v[n] * exp(DIR * 2*pi*i * n*k / R),   DIR = -1 (forward) or +1 (inverse, unnormalised).
// R is factored at compile time (4, then 2, then 3); twiddles are constexpr so that ptxas sees immediates
// (FFMA with an immediate operand issues at twice the rate of the 3-register form on Blackwell).
#pragma once
#include <cuda_runtime.h>
#include <utility>

#ifdef __CUDACC__
#define PTYB_HD __host__ __device__ __forceinline__
#define PTYB_CE __host__ __device__ constexpr
#else
#define PTYB_HD inline
#define PTYB_CE constexpr
#endif

namespace ptyb {

// ---- complex helpers on float2 -----------------------------------------------------------------
PTYB_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
PTYB_HD float2 cmulc(float2 a, float2 b) { /* a * conj(b) */ return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
PTYB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
PTYB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
PTYB_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
PTYB_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
PTYB_HD float cabs2(float2 a) { return a.x * a.x + a.y * a.y; }

// ---- constexpr cos/sin of 2*pi*k/r: exact quadrant reduction on integers, Taylor on |phi| <= pi/4 ----
constexpr double kPi = 3.14159265358979323846264338327950288;
PTYB_CE double taylor_sin(double x) {
    double x2 = x * x, t = x, s = x;
    for (int i = 1; i < 14; ++i) { t *= -x2 / double((2 * i) * (2 * i + 1)); s += t; }
    return s;
}
PTYB_CE double taylor_cos(double x) {
    double x2 = x * x, t = 1.0, s = 1.0;
    for (int i = 1; i < 14; ++i) { t *= -x2 / double((2 * i - 1) * (2 * i)); s += t; }
    return s;
}
struct cd { double c, s; };
PTYB_CE cd cs2pi(long long k, long long r) {
    k %= r;
    if (k < 0) k += r;
    long long k4 = 4 * k;
    int quad = int(k4 / r);
    long long rem = k4 - quad * r;                 // angle = quad*pi/2 + (pi/2)*rem/r
    double c = 1.0, s = 0.0;
    if (rem != 0) {
        if (2 * rem <= r) { double phi = (kPi / 2) * double(rem) / double(r); c = taylor_cos(phi); s = taylor_sin(phi); }
        else { double phi = (kPi / 2) * double(r - rem) / double(r); c = taylor_sin(phi); s = taylor_cos(phi); }
    }
    switch (quad) {
        case 0: return cd{c, s};
        case 1: return cd{-s, c};
        case 2: return cd{-c, -s};
        default: return cd{s, -c};
    }
}

// multiply by exp(DIR * i*pi/2) = DIR*i
template <int DIR> PTYB_HD float2 mul_i(float2 a) { return DIR > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// a * exp(DIR * 2*pi*i * K/R), K and R compile-time
template <int K, int R, int DIR> PTYB_HD float2 twmul(float2 a) {
    constexpr int k = ((K % R) + R) % R;
    if constexpr (k == 0) return a;
    else if constexpr (4 * k == R) return mul_i<DIR>(a);
    else if constexpr (2 * k == R) return make_float2(-a.x, -a.y);
    else if constexpr (4 * k == 3 * R) return mul_i<-DIR>(a);
    else if constexpr (8 * k == R || 8 * k == 3 * R || 8 * k == 5 * R || 8 * k == 7 * R) {
        constexpr float h = 0.70710678118654752440f;
        constexpr cd w = cs2pi(k, R);
        constexpr float sc = w.c > 0 ? 1.f : -1.f, ss = (DIR * w.s) > 0 ? 1.f : -1.f;
        // (x + i y) * h * (sc + i ss)
        return make_float2((a.x * sc - a.y * ss) * h, (a.x * ss + a.y * sc) * h);
    } else {
        constexpr cd w = cs2pi(k, R);
        constexpr float c = float(w.c), s = float(DIR * w.s);
        return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
    }
}

PTYB_CE int first_factor(int R) {
    return (R % 4 == 0 && R > 4) ? 4 : (R % 2 == 0 && R > 2) ? 2 : (R % 3 == 0 && R > 3) ? 3 : R;
}

// FOLD: fold the inter-stage twiddles into FMA butterflies (bfly_tw below): ~9 % fewer instructions per DFT32, slightly larger
// rounding error (the fused, issue-bound kernels use it; the HBM-bound row/column passes keep multiply-then-add, FOLD = false)
template <int R, int DIR, bool FOLD = true> struct Dft;

template <int DIR, bool FOLD> struct Dft<1, DIR, FOLD> { PTYB_HD static void run(float2*) {} };

template <int DIR, bool FOLD> struct Dft<2, DIR, FOLD> {
    PTYB_HD static void run(float2* v) {
        float2 t = v[0];
        v[0] = cadd(t, v[1]);
        v[1] = csub(t, v[1]);
    }
};

template <int DIR, bool FOLD> struct Dft<3, DIR, FOLD> {
    PTYB_HD static void run(float2* v) {
        constexpr float h = 0.86602540378443864676f * float(DIR);
        float2 t = cadd(v[1], v[2]), u = csub(v[1], v[2]);
        float2 m = make_float2(v[0].x - 0.5f * t.x, v[0].y - 0.5f * t.y);
        float2 iu = make_float2(-h * u.y, h * u.x);                 // DIR*i*(sqrt3/2)*u
        v[0] = cadd(v[0], t);
        v[1] = cadd(m, iu);
        v[2] = csub(m, iu);
    }
};

template <int DIR, bool FOLD> struct Dft<4, DIR, FOLD> {
    PTYB_HD static void run(float2* v) {
        float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
        float2 a2 = cadd(v[1], v[3]), a3 = mul_i<DIR>(csub(v[1], v[3]));
        v[0] = cadd(a0, a2);
        v[2] = csub(a0, a2);
        v[1] = cadd(a1, a3);
        v[3] = csub(a1, a3);
    }
};

// Butterfly with the twiddle FOLDED into the FMAs (Linzer-Feig):  op = p + w q,  om = p - w q,  w = exp(DIR * 2*pi*i * K/R).
// With w = c (1 + i t), t = tan (or w = s (cot + i) when |s| > |c|): u = (1 + i t) q costs 2 FFMA and p +- c u costs 4, six
// instructions where "complex multiply, then add and subtract" costs eight; t and c are compile-time immediates.
template <int K, int R, int DIR> PTYB_HD void bfly_tw(float2 p, float2 q, float2& op, float2& om) {
    constexpr int k = ((K % R) + R) % R;
    if constexpr (k == 0) { op = cadd(p, q); om = csub(p, q); }
    else if constexpr (2 * k == R) { op = csub(p, q); om = cadd(p, q); }
    else if constexpr (4 * k == R) { const float2 t = mul_i<DIR>(q); op = cadd(p, t); om = csub(p, t); }
    else if constexpr (4 * k == 3 * R) { const float2 t = mul_i<DIR>(q); op = csub(p, t); om = cadd(p, t); }
    else {
        constexpr cd w = cs2pi(k, R);
        constexpr double c = w.c, s = DIR * w.s;
        if constexpr ((c < 0 ? -c : c) >= (s < 0 ? -s : s)) {
            constexpr float t = float(s / c), cf = float(c);
            const float2 u = make_float2(fmaf(-t, q.y, q.x), fmaf(t, q.x, q.y));
            op = make_float2(fmaf(cf, u.x, p.x), fmaf(cf, u.y, p.y));
            om = make_float2(fmaf(-cf, u.x, p.x), fmaf(-cf, u.y, p.y));
        } else {
            constexpr float t = float(c / s), sf = float(s);
            const float2 u = make_float2(fmaf(t, q.x, -q.y), fmaf(t, q.y, q.x));
            op = make_float2(fmaf(sf, u.x, p.x), fmaf(sf, u.y, p.y));
            om = make_float2(fmaf(-sf, u.x, p.x), fmaf(-sf, u.y, p.y));
        }
    }
}

// second stage of the composite transform for one output residue KB: DFT_A over t[a] * W_R^{a KB}, twiddles folded
template <int R, int DIR, bool FOLD, int A, int B, int KB> PTYB_HD void second_stage(float2 (*y)[B], float2* v) {
    float2 t[A];
#pragma unroll
    for (int a = 0; a < A; ++a) t[a] = y[a][KB];
    if constexpr (!FOLD) {
        if constexpr (A > 1) t[1] = twmul<KB, R, DIR>(t[1]);
        if constexpr (A > 2) t[2] = twmul<2 * KB, R, DIR>(t[2]);
        if constexpr (A > 3) t[3] = twmul<3 * KB, R, DIR>(t[3]);
        Dft<A, DIR, FOLD>::run(t);
    } else if constexpr (A == 4) {
        // s_a = W^{a KB} t_a, W^{3 KB} = W^{KB} W^{2 KB}:  X0,2 = (t0 + W^2KB t2) +- W^KB (t1 + W^2KB t3),
        //                                                  X1,3 = (t0 - W^2KB t2) +- (DIR i) W^KB (t1 - W^2KB t3)
        float2 a0, a1, b0, b1;
        bfly_tw<2 * KB, R, DIR>(t[0], t[2], a0, a1);
        bfly_tw<2 * KB, R, DIR>(t[1], t[3], b0, b1);
        bfly_tw<KB, R, DIR>(a0, b0, t[0], t[2]);
        bfly_tw<KB + R / 4, R, DIR>(a1, b1, t[1], t[3]);
    } else if constexpr (A == 2) {
        bfly_tw<KB, R, DIR>(t[0], t[1], t[0], t[1]);
    } else {
        static_assert(A == 3, "radix");
        t[1] = twmul<KB, R, DIR>(t[1]);
        t[2] = twmul<2 * KB, R, DIR>(t[2]);
        Dft<A, DIR, FOLD>::run(t);
    }
#pragma unroll
    for (int ka = 0; ka < A; ++ka) v[KB + B * ka] = t[ka];
}
template <int R, int DIR, bool FOLD, int A, int B, int... KBs>
PTYB_HD void second_stages(float2 (*y)[B], float2* v, std::integer_sequence<int, KBs...>) {
    (second_stage<R, DIR, FOLD, A, B, KBs>(y, v), ...);
}

// composite R = A*B:  X[kb + B*ka] = sum_a W_A^{a ka} W_R^{a kb} sum_b W_B^{b kb} x[a + A b]
template <int R, int DIR, bool FOLD> struct Dft {
    static constexpr int A = first_factor(R), B = R / A;
    static_assert(A != R, "prime radix not implemented");
    PTYB_HD static void run(float2* v) {
        float2 y[A][B];
#pragma unroll
        for (int a = 0; a < A; ++a)
#pragma unroll
            for (int b = 0; b < B; ++b) y[a][b] = v[a + A * b];
#pragma unroll
        for (int a = 0; a < A; ++a) Dft<B, DIR, FOLD>::run(y[a]);
        second_stages<R, DIR, FOLD, A, B>(y, v, std::make_integer_sequence<int, B>{});
    }
};

}  // namespace ptyb
