// C ABI of libptyrad_b200.so (see include/ptyrad_b200.h).  Host-side sequencing of the kernels; no allocation,
// no synchronisation: everything is enqueued on the caller's stream.
#include "../../include/ptyrad_b200.h"
#include "general_kernels.cuh"
#include "fused128.cuh"
#include "fused64.cuh"
#include "grouping.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>

using namespace ptyb;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};   // kernels launched by this library (bench.py reports it as gpu_launches)

// optional per-section device timing (bench.py roofline): event pairs around the multislice forward / adjoint sections
constexpr int kMaxEv = 256;
struct Timing {
    bool on = false;
    cudaEvent_t ev[2][kMaxEv][2];
    int n[2] = {0, 0};
    bool created = false;
} g_tm;
std::mutex g_tm_mu;
void tm_mark(int section, int which, cudaStream_t st) {
    if (!g_tm.on) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;   // no events inside a graph capture
    std::lock_guard<std::mutex> lk(g_tm_mu);
    if (!g_tm.created) {
        for (int s = 0; s < 2; ++s) for (int i = 0; i < kMaxEv; ++i) { cudaEventCreate(&g_tm.ev[s][i][0]); cudaEventCreate(&g_tm.ev[s][i][1]); }
        g_tm.created = true;
    }
    if (g_tm.n[section] >= kMaxEv) return;
    cudaEventRecord(g_tm.ev[section][g_tm.n[section]][which], st);
    if (which == 1) ++g_tm.n[section];
}

// Independent short kernels of one call (object polar form | propagator transpose + permute | probe-spectrum FFT + permute before
// the multislice kernels; probe-gradient inverse FFT | object polar backward after them) run as parallel branches: a step is a chain
// of ~30 launches of which two matter, and every link of a serial chain costs a few microseconds of launch latency even inside a
// CUDA graph.  The branches use two internal non-blocking streams that are forked from and joined back into the caller's stream
// with events inside the same call, so the caller still sees one stream (and a stream capture sees a fork/join subgraph).
struct SidePool {
    cudaStream_t s[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
    bool ok = false;
    int dev = -1;                                          // the device the streams live on (one process per GPU is the intended use)
    bool init() {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) return false;
        if (ok) return cur == dev;                         // a call on another device of the same process stays serial
        dev = cur;
        for (int i = 0; i < 2; ++i) {
            if (cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess) return false;
        ok = true;
        return true;
    }
} g_side;
std::mutex g_side_mu;
static const bool g_no_branches = [] { const char* e = getenv("PTYB200_NO_BRANCHES"); return e && atoi(e) != 0; }();
// fork: both side streams wait for everything enqueued on `st` so far; returns false (caller stays serial) if unavailable
bool side_fork(cudaStream_t st, cudaStream_t out[2]) {
    out[0] = out[1] = st;
    if (g_no_branches) return false;
    std::lock_guard<std::mutex> lk(g_side_mu);
    if (!g_side.init()) { cudaGetLastError(); return false; }
    if (cudaEventRecord(g_side.fork, st) != cudaSuccess) { cudaGetLastError(); return false; }
    for (int i = 0; i < 2; ++i)
        if (cudaStreamWaitEvent(g_side.s[i], g_side.fork, 0) != cudaSuccess) { cudaGetLastError(); return false; }
    out[0] = g_side.s[0]; out[1] = g_side.s[1];
    return true;
}
// join: `st` waits for both side streams
int side_join(cudaStream_t st, bool forked) {
    if (!forked) return 0;
    std::lock_guard<std::mutex> lk(g_side_mu);
    for (int i = 0; i < 2; ++i) {
        if (cudaEventRecord(g_side.join[i], g_side.s[i]) != cudaSuccess) return 1;
        if (cudaStreamWaitEvent(st, g_side.join[i], 0) != cudaSuccess) return 1;
    }
    return 0;
}

int fail(const char* what, cudaError_t e, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    g_err = buf;
    return 1;
}
int fail_msg(const std::string& m) { g_err = m; return 2; }

#define CK(call)                                                        \
    do {                                                                \
        cudaError_t e_ = (call);                                        \
        if (e_ != cudaSuccess) return fail(#call, e_, __FILE__, __LINE__); \
    } while (0)
#define CKL() do { ++g_launches; CK(cudaGetLastError()); } while (0)

size_t al(size_t x) { return (x + 255) & ~size_t(255); }

struct Workspace {
    float2 *O, *gO, *PhatT, *tmpP, *gPhatT, *HT, *wvec, *tvec, *stash, *phis, *G1, *G2, *farT;
    float* gprop;
    unsigned char* fused;      // scratch owned by the fused path
    size_t total;
};

bool supported_N(int N) { return N == 16 || N == 32 || N == 48 || N == 64 || N == 96 || N == 128 || N == 192 || N == 256; }
// register-resident on-chip kernels exist for N = 128 (fused128.cuh) and N = 64 (fused64.cuh)
bool fused_covers(const ptyb200_cfg& c) { return fused128::covers(c) || fused64::covers(c); }
bool use_fused(const ptyb200_cfg& c) { return c.path != PTYB200_PATH_GENERAL && fused_covers(c); }
size_t fused_scratch_bytes(const ptyb200_cfg& c, int B) { return fused64::covers(c) ? fused64::scratch_bytes(c, B) : fused128::scratch_bytes(c, B); }

// General path: optionally the slice sequence runs on CHUNKS of the batch (pass buffers G1/G2 sized for one chunk, so that they can
// stay L2-resident from the kernel that writes a tile to the kernel that reads it) with the probe modes split into per-CTA groups
// (to keep the launches wide when the chunk is small).  Measured on B200 (profiles/r01/README.md, "chunked general path"): the
// per-slice kernels are bound by shared-memory bandwidth and issue, not by HBM, so chunking only adds launch tails (C4: 74 ms
// unchunked, 98-137 ms chunked) -- the default is therefore ONE chunk and all probe modes per CTA.  cfg.reserved[2] / [3] (or the
// environment variables PTYB200_GEN_CHUNK / PTYB200_GEN_PG, for sweeps) select a cut; tests exercise ragged cuts.
struct GenPlan { int chunk, pg, groups; };
GenPlan gen_plan(const ptyb200_cfg& c, int B) {
    GenPlan g;
    // the environment is read ONCE per process: a change between ptyb200_workspace_bytes and the launch must not move the cut
    static const int env_chunk = [] { const char* e = getenv("PTYB200_GEN_CHUNK"); return e ? atoi(e) : 0; }();
    static const int env_pg = [] { const char* e = getenv("PTYB200_GEN_PG"); return e ? atoi(e) : 0; }();
    int chunk = B, pg = c.P;
    if (env_chunk > 0) chunk = env_chunk;
    if (c.reserved[2] > 0) chunk = c.reserved[2];
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
    if (env_pg > 0) pg = env_pg;
    if (c.reserved[3] > 0) pg = c.reserved[3];
    if (pg < 1) pg = 1;
    if (pg > c.P) pg = c.P;
    g.chunk = chunk; g.pg = pg; g.groups = (c.P + pg - 1) / pg;
    return g;
}
constexpr int kProbeCtas = 4 * 148 * 4;   // probe pass: ~4 waves of 148 SMs x 4 resident CTAs, so that no CTA loops over more than B/7 samples

// B = samples of this call; the layout follows the batch CAPACITY the workspace was sized for (cfg.reserved[0], chunked steps), so
// that the gradient accumulators keep their place when the last chunk is smaller
Workspace carve(const ptyb200_cfg& c, int B_call, void* base) {
    const int B = c.reserved[0] > B_call ? c.reserved[0] : B_call;
    Workspace w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return reinterpret_cast<unsigned char*>(base) + o; };
    const size_t NN = (size_t)c.N * c.N, obj = (size_t)((c.reserved[1] & 1) ? B : 1) * c.M * c.Z * c.Noy * c.Nox, tiles = (size_t)B * c.P * c.M;
    w.O = (float2*)take(obj * 8);
    w.gO = (float2*)take(obj * 8);
    w.PhatT = (float2*)take(c.P * NN * 8);
    w.tmpP = (float2*)take(c.P * NN * 8);
    w.gPhatT = (float2*)take(c.P * NN * 8);
    w.HT = (float2*)take(NN * 8);
    w.wvec = (float2*)take((size_t)B * 2 * c.N * 8);
    w.tvec = (float2*)take((size_t)B * 2 * c.N * 8);
    w.gprop = (float*)take((size_t)B * 3 * 4);
    w.stash = (float2*)take(tiles * c.Z * NN * 8);
    w.phis = (float2*)take(c.stash_fourier ? tiles * (c.Z > 1 ? c.Z - 1 : 0) * NN * 8 : 0);
    const size_t ctiles = use_fused(c) ? 0 : (size_t)gen_plan(c, B).chunk * c.P * c.M;    // pass buffers: one chunk of the batch
    w.G1 = (float2*)take(ctiles * NN * 8);
    w.G2 = (float2*)take(ctiles * NN * 8);
    w.farT = (float2*)take(use_fused(c) ? 0 : tiles * NN * 8);
    w.fused = take(use_fused(c) ? fused_scratch_bytes(c, B) : 0);
    (void)B_call;
    w.total = off;
    return w;
}

int check_cfg(const ptyb200_cfg* c, int B) {
    if (!c) return fail_msg("cfg is NULL");
    if (!supported_N(c->N)) return fail_msg("unsupported N=" + std::to_string(c->N) + " (supported: 16,32,48,64,96,128,192,256)");
    if (c->P < 1 || c->M < 1 || c->Z < 1 || B < 1) return fail_msg("P, M, Z and B must be >= 1");
    if (c->Noy < c->N || c->Nox < c->N) return fail_msg("object canvas smaller than the probe");
    if (c->tilt_mode < 0 || c->tilt_mode > 2) return fail_msg("tilt_mode must be 0, 1 or 2");
    if ((c->reserved[1] & 1) && (c->Noy != c->N || c->Nox != c->N)) return fail_msg("patch mode needs Noy == Nox == N");
    return 0;
}

#define DISPATCH_N(N_, ...)                                                              \
    switch (N_) {                                                                         \
        case 16:  { using F = RowFFT<4, 4>;   __VA_ARGS__; } break;                       \
        case 32:  { using F = RowFFT<8, 4>;   __VA_ARGS__; } break;                       \
        case 48:  { using F = RowFFT<8, 6>;   __VA_ARGS__; } break;                       \
        case 64:  { using F = RowFFT<8, 8>;   __VA_ARGS__; } break;                       \
        case 96:  { using F = RowFFT<12, 8>;  __VA_ARGS__; } break;                       \
        case 128: { using F = RowFFT<16, 8>;  __VA_ARGS__; } break;                       \
        case 192: { using F = RowFFT<16, 12>; __VA_ARGS__; } break;                       \
        case 256: { using F = RowFFT<16, 16>; __VA_ARGS__; } break;                       \
        default: return fail_msg("unsupported N");                                        \
    }

template <class K> cudaError_t prep(K kern, size_t bytes) {
    return bytes > 48 * 1024 ? cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) : cudaSuccess;
}

#define LAUNCH(kern, grid, stream, ...)                                   \
    do {                                                                  \
        const size_t sm_ = Slab<F>::smem_bytes();                         \
        CK(prep(kern, sm_));                                              \
        kern<<<grid, NT, sm_, stream>>>(__VA_ARGS__);                     \
        CKL();                                                            \
    } while (0)

template <class F> int fft2_tiles(const float2* in, float2* tmp, float2* out, int count, int dir, cudaStream_t st) {
    dim3 g(F::N / ROWS, count);
    if (dir < 0) {
        LAUNCH((k_pass<F, -1, true>), g, st, in, tmp, 1.0f);
        LAUNCH((k_pass<F, -1, false>), g, st, tmp, out, 1.0f);
    } else {
        LAUNCH((k_pass<F, +1, true>), g, st, in, tmp, 1.0f);
        LAUNCH((k_pass<F, +1, false>), g, st, tmp, out, 1.0f);
    }
    return 0;
}

FwdArgs make_fwd_args(const ptyb200_cfg& c, int B, const Workspace& w, const int64_t* idx, const int32_t* crop, const float* probe,
                      const float* occu, float* dp) {
    FwdArgs a;
    a.d = Dims{c.N, c.P, c.M, c.Z, c.Noy, c.Nox, B, (c.reserved[1] & 1), 0, c.P};
    a.idx = idx; a.crop = crop; a.O = w.O; a.probe = (const float2*)probe; a.PhatT = w.PhatT; a.HT = w.HT;
    a.wvec = c.shift_probes ? w.wvec : nullptr;
    a.tvec = c.tilt_mode ? w.tvec : nullptr;
    a.occu = occu; a.stash = w.stash; a.phis = c.stash_fourier && c.Z > 1 ? w.phis : nullptr;
    a.G1 = w.G1; a.G2 = w.G2; a.farT = w.farT; a.dp = dp; a.eps = c.eps;
    memset(&a.lf, 0, sizeof a.lf);
    return a;
}

// shared setup: complex object, transposed propagator, probe spectrum, per-sample ramps
template <class F> int setup_common(const ptyb200_cfg& c, int B, const Workspace& w, const int64_t* idx, const float* obja,
                                    const float* objp, const float* probe, const float* shifts, const float* Hbase,
                                    const float* tilts, const float* dz, cudaStream_t st, cudaStream_t sH, cudaStream_t sP) {
    // st: complex object + per-sample ramps; sH: transposed propagator; sP: probe spectrum (the caller joins sH / sP into st)
    const size_t obj = (size_t)((c.reserved[1] & 1) ? B : 1) * c.M * c.Z * c.Noy * c.Nox;
    k_obj_polar<<<(unsigned)((obj + 255) / 256), 256, 0, st>>>(obja, objp, w.O, obj);
    CKL();
    dim3 tb(32, 8), tg((c.N + 31) / 32, (c.N + 31) / 32);
    k_transpose<<<tg, tb, 0, sH>>>((const float2*)Hbase, w.HT, c.N);
    CKL();
    if (c.shift_probes) {
        if (!shifts) return fail_msg("shift_probes set but shifts is NULL");
        if (int r = fft2_tiles<F>((const float2*)probe, w.tmpP, w.PhatT, c.P, -1, sP)) return r;
        k_shift_vectors<<<B, 128, 0, st>>>(shifts, idx, B, c.N, w.wvec);
        CKL();
    }
    if (c.tilt_mode) {
        if (!tilts || !dz) return fail_msg("tilt_mode set but tilts/dz is NULL");
        k_tilt_vectors<<<B, 128, 0, st>>>(tilts, c.tilt_mode, idx, B, c.N, c.dx, dz, w.tvec);
        CKL();
    }
    return 0;
}

template <class F> int forward_general(const ptyb200_cfg& c, int B, const Workspace& w, FwdArgs a, cudaStream_t st) {
    const int nb = c.N / ROWS;
    const GenPlan gp = gen_plan(c, B);
    a.d.pg = gp.pg;
    for (int b0 = 0; b0 < B; b0 += gp.chunk) {            // the whole slice sequence per chunk: G1/G2 never leave L2
        const int nbc = B - b0 < gp.chunk ? B - b0 : gp.chunk;
        a.d.b0 = b0;
        const dim3 grid(nb, c.M * gp.groups, nbc);
        if (c.shift_probes) LAUNCH((k_init_shift<F>), dim3(nb, gp.groups, nbc), st, a);
        for (int z = 0; z < c.Z; ++z) {
            const int src_mode = (z == 0 && !c.shift_probes) ? 1 : 0;
            LAUNCH((k_fwd_da<F>), grid, st, a, z, src_mode, z == c.Z - 1 ? 1 : 0);
            if (z < c.Z - 1) LAUNCH((k_fwd_bc<F>), grid, st, a, z);
        }
    }
    a.d.b0 = 0;
    LAUNCH((k_fwd_final<F>), dim3(nb, B), st, a);        // far-field tiles of the whole batch
    return 0;
}

template <class F> int backward_general(const ptyb200_cfg& c, int B, const Workspace& w, BwdArgs a, float* g_probe, cudaStream_t st, int acc = 0) {
    const int nb = c.N / ROWS;
    const bool want_p = a.need_probe || a.need_shift;
    const GenPlan gp = gen_plan(c, B);
    a.f.d.pg = gp.pg;
    for (int b0 = 0; b0 < B; b0 += gp.chunk) {
        const int nbc = B - b0 < gp.chunk ? B - b0 : gp.chunk;
        a.f.d.b0 = b0;
        const dim3 grid(nb, c.M * gp.groups, nbc);
        LAUNCH((k_bwd_start<F>), grid, st, a);
        for (int z = c.Z - 1; z >= 0; --z) {
            int out_mode = 0;
            if (z == 0) out_mode = want_p ? (c.shift_probes ? 0 : 1) : 2;
            LAUNCH((k_bwd_da<F>), grid, st, a, z, out_mode);
            if (z > 0) {
                if (a.need_prop) LAUNCH((k_bwd_bc<F, true>), grid, st, a, z);
                else LAUNCH((k_bwd_bc<F, false>), grid, st, a, z);
            }
        }
        if (want_p) {                                     // gpsi_0 of this chunk is in G1
            if (c.shift_probes) {
                int per = nb * c.P;
                int nsub = (kProbeCtas + per - 1) / per;
                if (nsub > nbc) nsub = nbc;
                if (nsub < 1) nsub = 1;
                int bsub = (nbc + nsub - 1) / nsub;
                nsub = (nbc + bsub - 1) / bsub;
                LAUNCH((k_bwd_probe<F>), dim3(nb, c.P, nsub), st, a, nbc, bsub);
            } else if (a.need_probe) {
                k_bwd_probe_noshift<<<dim3((c.N * c.N + 255) / 256, c.P), 256, 0, st>>>(a.f.d, nbc, (b0 == 0 && !(acc & PTYB200_ACC_KEEP_GRADS)) ? 1 : 0, w.G1, (float2*)g_probe);
                CKL();
            }
        }
    }
    a.f.d.b0 = 0;
    if (want_p && c.shift_probes && a.need_probe && !(acc & PTYB200_ACC_NO_FINISH))
        if (int r = fft2_tiles<F>(w.gPhatT, w.tmpP, (float2*)g_probe, c.P, +1, st)) return r;
    return 0;
}

// kernel-side view of the measurements from the ABI's description; returns an error text or null
const char* make_meas_view(const ptyb200_cfg& c, const ptyb200_meas_cfg* m, const float* meas_all, const float* padded, const void* a1,
                           const void* a2, MeasView* out) {
    MeasView v;
    memset(&v, 0, sizeof v);
    v.meas = meas_all;
    v.Hs = v.Ws = c.N;
    v.ry = v.rx = v.scale = 1.f;
    if (m) {
        if (m->Hs < 1 || m->Ws < 1) return "meas_cfg: stored pattern size must be positive";
        v.Hs = m->Hs; v.Ws = m->Ws;
        int H = v.Hs, W = v.Ws;
        if (m->Hp > 0 || m->Wp > 0) {
            if (!padded) return "meas_cfg: padded canvas size given but meas_padded is NULL";
            if (m->h1 < 0 || m->w1 < 0 || m->h2 > m->Hp || m->w2 > m->Wp || m->h2 - m->h1 != m->Hs || m->w2 - m->w1 != m->Ws)
                return "meas_cfg: paste window [h1:h2, w1:w2] must lie inside the canvas and match the stored pattern size";
            v.padded = padded; v.Hp = m->Hp; v.Wp = m->Wp; v.h1 = m->h1; v.h2 = m->h2; v.w1 = m->w1; v.w2 = m->w2;
            H = v.Hp; W = v.Wp;
        }
        const bool rs = (m->scale_y > 0.f && m->scale_y != 1.f) || (m->scale_x > 0.f && m->scale_x != 1.f);
        if (rs) {
            const float sy = m->scale_y > 0.f ? m->scale_y : 1.f, sx = m->scale_x > 0.f ? m->scale_x : 1.f;
            v.resample = 1; v.ry = 1.f / sy; v.rx = 1.f / sx; v.scale = 1.f / (sy * sx);
            H = (int)floor((double)H * (double)sy); W = (int)floor((double)W * (double)sx);
        }
        if (H != c.N || W != c.N) return "meas_cfg: padded / resampled pattern size does not equal cfg.N";
    }
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    v.vec = (!v.padded && !v.resample && al16(meas_all) && al16(a1) && al16(a2) && (c.N * c.N) % 4 == 0) ? 1 : 0;
    *out = v;
    return nullptr;
}

LossK make_lossk(const ptyb200_loss_cfg& l) {
    LossK k;
    k.s_on = l.single_state; k.s_w = l.single_weight; k.s_p = l.single_pow;
    k.p_on = l.poissn_state; k.p_w = l.poissn_weight; k.p_p = l.poissn_pow; k.p_eps = l.poissn_eps;
    k.b_on = l.pacbed_state; k.b_w = l.pacbed_weight; k.b_p = l.pacbed_pow;
    return k;
}

}  // namespace

extern "C" {

int ptyb200_abi_version(void) { return PTYB200_ABI_VERSION; }
long long ptyb200_launch_count(void) { return g_launches.load(); }
void ptyb200_timing_enable(int on) { std::lock_guard<std::mutex> lk(g_tm_mu); g_tm.on = on != 0; g_tm.n[0] = g_tm.n[1] = 0; }
int ptyb200_timing_read(double* ms_forward, double* ms_backward, int* n_forward, int* n_backward) {
    std::lock_guard<std::mutex> lk(g_tm_mu);
    double acc[2] = {0, 0};
    for (int s = 0; s < 2; ++s)
        for (int i = 0; i < g_tm.n[s]; ++i) {
            float ms = 0.f;
            CK(cudaEventSynchronize(g_tm.ev[s][i][1]));
            CK(cudaEventElapsedTime(&ms, g_tm.ev[s][i][0], g_tm.ev[s][i][1]));
            acc[s] += ms;
        }
    if (ms_forward) *ms_forward = acc[0];
    if (ms_backward) *ms_backward = acc[1];
    if (n_forward) *n_forward = g_tm.n[0];
    if (n_backward) *n_backward = g_tm.n[1];
    g_tm.n[0] = g_tm.n[1] = 0;
    return 0;
}
const char* ptyb200_last_error(void) { return g_err.c_str(); }

size_t ptyb200_workspace_bytes(const ptyb200_cfg* cfg, int32_t B) {
    if (check_cfg(cfg, B)) return 0;
    return carve(*cfg, B, nullptr).total;
}

int ptyb200_propagator(const ptyb200_cfg* c, const float* dz, float* H_out, ptyb200_stream s) {
    if (!c || !dz || !H_out) return fail_msg("NULL argument");
    k_propagator<<<(c->N * c->N + 255) / 256, 256, 0, (cudaStream_t)s>>>(c->N, c->dx, c->lambd, dz, (float2*)H_out);
    CKL();
    return 0;
}

int ptyb200_gather_patches(const ptyb200_cfg* c, const int64_t* idx, int32_t B, const float* obja, const float* objp,
                           const int32_t* crop_pos, float* patches_out, ptyb200_stream s) {
    if (!c || !idx || !obja || !objp || !crop_pos || !patches_out) return fail_msg("NULL argument");
    Dims d{c->N, c->P, c->M, c->Z, c->Noy, c->Nox, B, 0};
    dim3 g((c->N * c->N + 1023) / 1024, c->M * c->Z, B);
    k_gather_patches<<<g, 256, 0, (cudaStream_t)s>>>(d, idx, obja, objp, crop_pos, patches_out);
    CKL();
    return 0;
}

static int forward_impl(const ptyb200_cfg* c, const int64_t* idx, int32_t B, const float* obja, const float* objp,
                        const int32_t* crop_pos, const float* probe, const float* shifts, const float* Hbase,
                        const float* tilts, const float* dz, const float* occu, float* dp_out, void* workspace,
                        const LossFuse* lf, ptyb200_stream s) {
    if (int r = check_cfg(c, B)) return r;
    if (!idx || !obja || !objp || (!crop_pos && !(c->reserved[1] & 1)) || !probe || !Hbase || !occu || !dp_out || !workspace) return fail_msg("NULL argument");
    cudaStream_t st = (cudaStream_t)s;
    Workspace w = carve(*c, B, workspace);
    FwdArgs a = make_fwd_args(*c, B, w, idx, crop_pos, probe, occu, dp_out);
    if (lf) a.lf = *lf;
    if (c->path == PTYB200_PATH_FUSED && !fused_covers(*c)) return fail_msg("fused path does not cover this configuration");
    cudaStream_t side[2];
    const bool forked = side_fork(st, side);
    DISPATCH_N(c->N, {
        if (int r = setup_common<F>(*c, B, w, idx, obja, objp, probe, shifts, Hbase, tilts, dz, st, side[0], side[1])) return r;
        if (use_fused(*c)) {
            // the fused paths permute the two tables on the branches that made them and join inside, just before the wave kernel
            auto join = [&]() -> int { if (side_join(st, forked)) return fail_msg("side-stream join failed"); tm_mark(0, 0, st); return 0; };
            if (int r = fused64::covers(*c) ? fused64::forward(*c, B, a, obja, objp, w.fused, st, g_err, &g_launches, side[0], side[1], join)
                                            : fused128::forward(*c, B, a, obja, objp, w.fused, st, g_err, &g_launches, side[0], side[1], join)) return r;
        } else {
            if (side_join(st, forked)) return fail_msg("side-stream join failed");
            tm_mark(0, 0, st);
            if (int r = forward_general<F>(*c, B, w, a, st)) return r;
        }
        tm_mark(0, 1, st);
    });
    return 0;
}

int ptyb200_forward(const ptyb200_cfg* c, const int64_t* idx, int32_t B, const float* obja, const float* objp,
                    const int32_t* crop_pos, const float* probe, const float* shifts, const float* Hbase,
                    const float* tilts, const float* dz, const float* occu, float* dp_out, void* workspace,
                    ptyb200_stream s) {
    return forward_impl(c, idx, B, obja, objp, crop_pos, probe, shifts, Hbase, tilts, dz, occu, dp_out, workspace, nullptr, s);
}

int ptyb200_forward_loss(const ptyb200_cfg* c, const int64_t* idx, int32_t B, const float* obja, const float* objp,
                         const int32_t* crop_pos, const float* probe, const float* shifts, const float* Hbase,
                         const float* tilts, const float* dz, const float* occu, float* dp_out, void* workspace,
                         const ptyb200_loss_cfg* lc, const float* meas_all, const int64_t* meas_rows, const ptyb200_meas_cfg* mcfg,
                         const float* meas_padded, float* losses3, double* stats, float* pac, ptyb200_stream s) {
    if (!c || !lc || !meas_all || !losses3 || !stats || !dp_out) return fail_msg("NULL argument");
    if (lc->pacbed_state && !pac) return fail_msg("pacbed needs pacbed_scratch");
    cudaStream_t st = (cudaStream_t)s;
    LossFuse lf;
    memset(&lf, 0, sizeof lf);
    lf.on = 1;
    lf.k = make_lossk(*lc);
    if (const char* e = make_meas_view(*c, mcfg, meas_all, meas_padded, dp_out, dp_out, &lf.mv)) return fail_msg(e);
    lf.stats = stats; lf.pac = pac; lf.rows = meas_rows ? meas_rows : idx;
    if (!(c->reserved[4] & PTYB200_ACC_KEEP_STATS)) {
        CK(cudaMemsetAsync(stats, 0, 8 * sizeof(double), st));
        if (lf.k.b_on) CK(cudaMemsetAsync(pac, 0, (size_t)2 * c->N * c->N * 4, st));
    }
    if (int r = forward_impl(c, idx, B, obja, objp, crop_pos, probe, shifts, Hbase, tilts, dz, occu, dp_out, workspace, &lf, s)) return r;
    if (!(c->reserved[4] & PTYB200_ACC_NO_LOSS_FINAL)) {
        k_loss_final<<<1, 256, 0, st>>>(lf.k, B, c->N, stats, pac, losses3);
        CKL();
    }
    return 0;
}

int ptyb200_backward(const ptyb200_cfg* c, const int64_t* idx, int32_t B, const float* obja, const float* objp,
                     const int32_t* crop_pos, const float* probe, const float* shifts, const float* Hbase,
                     const float* tilts, const float* dz, const float* occu, const float* G, void* workspace,
                     float* g_obja, float* g_objp, float* g_probe, float* g_shifts, float* g_tilts, float* g_dz,
                     uint32_t need_mask, ptyb200_stream s) {
    if (int r = check_cfg(c, B)) return r;
    if (!idx || !obja || !objp || (!crop_pos && !(c->reserved[1] & 1)) || !probe || !Hbase || !occu || !G || !workspace) return fail_msg("NULL argument");
    cudaStream_t st = (cudaStream_t)s;
    Workspace w = carve(*c, B, workspace);
    BwdArgs a;
    a.f = make_fwd_args(*c, B, w, idx, crop_pos, probe, occu, nullptr);
    a.G = G; a.gO = w.gO; a.gPhatT = w.gPhatT; a.gprop = w.gprop; a.gshift = g_shifts;
    a.dx = c->dx; a.k0 = 6.283185307179586f / c->lambd;
    a.need_obj = (need_mask & PTYB200_NEED_OBJ) ? 1 : 0;
    a.need_probe = (need_mask & PTYB200_NEED_PROBE) ? 1 : 0;
    a.need_shift = ((need_mask & PTYB200_NEED_SHIFTS) && c->shift_probes) ? 1 : 0;
    const bool need_t = (need_mask & PTYB200_NEED_TILTS) != 0, need_dz = (need_mask & PTYB200_NEED_DZ) != 0;
    a.need_prop = ((need_t || need_dz) && c->Z > 1) ? 1 : 0;
    if (a.need_obj && (!g_obja || !g_objp)) return fail_msg("g_obja/g_objp is NULL");
    if (a.need_probe && !g_probe) return fail_msg("g_probe is NULL");
    if ((need_mask & PTYB200_NEED_SHIFTS) && !g_shifts) return fail_msg("g_shifts is NULL");
    if (need_t && (!g_tilts || !c->tilt_mode)) return fail_msg("tilt gradient requested without tilts");
    if (need_dz && !g_dz) return fail_msg("g_dz is NULL");
    if (a.need_prop && !c->stash_fourier) return fail_msg("tilt/thickness gradients need cfg.stash_fourier = 1 in the forward");
    // the kernels write gradients with 8- and 16-byte vector accesses
    auto misaligned = [](const void* p, size_t al) { return p && (reinterpret_cast<uintptr_t>(p) & (al - 1)) != 0; };
    if (misaligned(g_obja, 16) || misaligned(g_objp, 16) || misaligned(g_probe, 16) || misaligned(g_shifts, 8) || misaligned(g_tilts, 8))
        return fail_msg("gradient buffers must be 16-byte aligned (g_shifts / g_tilts: 8-byte)");
    const size_t obj = (size_t)((c->reserved[1] & 1) ? B : 1) * c->M * c->Z * c->Noy * c->Nox;
    const int acc = c->reserved[4];
    if (acc & PTYB200_ACC_NO_FINISH) {                    // (KEEP_GRADS alone = accumulators zeroed early by ptyb200_backward_zero: any configuration)
        if (a.need_prop) return fail_msg("chunked steps do not cover tilt / thickness gradients");
        if (c->reserved[1] & 1) return fail_msg("chunked steps do not cover patch mode");
    }
    if (!(acc & PTYB200_ACC_KEEP_GRADS)) {
        if (a.need_obj && !use_fused(*c)) CK(cudaMemsetAsync(w.gO, 0, obj * 8, st));
        if (a.need_probe && c->shift_probes) CK(cudaMemsetAsync(w.gPhatT, 0, (size_t)c->P * c->N * c->N * 8, st));
        if (need_mask & PTYB200_NEED_SHIFTS) CK(cudaMemsetAsync(g_shifts, 0, (size_t)c->Ntot * 2 * 4, st));
    }
    if (need_t || need_dz) CK(cudaMemsetAsync(w.gprop, 0, (size_t)B * 3 * 4, st));
    if (need_t) CK(cudaMemsetAsync(g_tilts, 0, (size_t)(c->tilt_mode == 2 ? c->Ntot : 1) * 2 * 4, st));
    if (need_dz) CK(cudaMemsetAsync(g_dz, 0, 4, st));
    if (c->path == PTYB200_PATH_FUSED && !fused_covers(*c)) return fail_msg("fused path does not cover this configuration");
    // after the adjoint kernel of the fused paths: probe-spectrum gradient (unpermute + inverse FFT) on a branch, object polar backward
    // on the caller's stream
    cudaStream_t sfin[2] = {st, st};
    bool fin_forked = false;
    auto fork_fin = [&]() -> cudaStream_t { tm_mark(1, 1, st); fin_forked = side_fork(st, sfin); return sfin[0]; };
    DISPATCH_N(c->N, {
        tm_mark(1, 0, st);
        if (use_fused(*c)) {
            if (int r = fused64::covers(*c)
                            ? fused64::backward(*c, B, a, obja, objp, g_obja, g_objp, w.fused, (float2*)g_probe, w.gPhatT, st, g_err, &g_launches, acc, nullptr, false, fork_fin)
                            : fused128::backward(*c, B, a, obja, objp, g_obja, g_objp, w.fused, (float2*)g_probe, w.gPhatT, st, g_err, &g_launches, acc, nullptr, false, fork_fin)) return r;
        }
        else if (int r = backward_general<F>(*c, B, w, a, g_probe, st, acc)) return r;
        if (!use_fused(*c)) tm_mark(1, 1, st);
        if (use_fused(*c) && a.need_probe && c->shift_probes && !(acc & PTYB200_ACC_NO_FINISH))
            if (int r = fft2_tiles<F>(w.gPhatT, w.tmpP, (float2*)g_probe, c->P, +1, sfin[0])) return r;
    });
    if (a.need_obj && !use_fused(*c) && !(acc & PTYB200_ACC_NO_FINISH)) {
        k_obj_finish<<<(unsigned)((obj + 255) / 256), 256, 0, st>>>(w.gO, obja, objp, g_obja, g_objp, obj, nullptr, (acc & PTYB200_ACC_ADD_OBJ) ? 1 : 0);
        CKL();
    }
    if (side_join(st, fin_forked)) return fail_msg("side-stream join failed");
    if (a.need_prop) {
        k_prop_finish<<<1, 256, 0, st>>>(w.gprop, tilts, c->tilt_mode, idx, B, dz, need_t ? g_tilts : nullptr, need_dz ? g_dz : nullptr);
        CKL();
    }
    return 0;
}

int ptyb200_gather_measurements(const ptyb200_cfg* c, const ptyb200_meas_cfg* mcfg, const float* meas_all, const float* meas_padded,
                                const int64_t* idx, int32_t B, float* out, ptyb200_stream s) {
    if (!c || !meas_all || !idx || !out || B < 1) return fail_msg("NULL argument");
    MeasView mv;
    if (const char* e = make_meas_view(*c, mcfg, meas_all, meas_padded, out, out, &mv)) return fail_msg(e);
    mv.vec = 0;
    unsigned chunks = (unsigned)((c->N * c->N + 256 * 8 - 1) / (256 * 8));
    k_meas_gather<<<dim3(chunks, B), 256, 0, (cudaStream_t)s>>>(mv, idx, c->N, out);
    CKL();
    return 0;
}

int ptyb200_loss_forward(const ptyb200_cfg* c, const ptyb200_loss_cfg* lc, const float* dp, const float* meas_all,
                         const int64_t* idx, int32_t B, float* losses3, double* stats, float* pac,
                         const ptyb200_meas_cfg* mcfg, const float* meas_padded, ptyb200_stream s) {
    if (!c || !lc || !dp || !meas_all || !idx || !losses3 || !stats) return fail_msg("NULL argument");
    if (lc->pacbed_state && !pac) return fail_msg("pacbed needs pacbed_scratch");
    cudaStream_t st = (cudaStream_t)s;
    LossK k = make_lossk(*lc);
    MeasView mv;
    if (const char* e = make_meas_view(*c, mcfg, meas_all, meas_padded, dp, dp, &mv)) return fail_msg(e);
    CK(cudaMemsetAsync(stats, 0, 8 * sizeof(double), st));
    if (k.b_on) CK(cudaMemsetAsync(pac, 0, (size_t)2 * c->N * c->N * 4, st));
    unsigned chunks = (unsigned)((c->N * c->N + 256 * 16 - 1) / (256 * 16));
    k_loss_partial<<<dim3(chunks, B), 256, 0, st>>>(k, dp, mv, idx, B, c->N, stats, pac);
    CKL();
    k_loss_final<<<1, 256, 0, st>>>(k, B, c->N, stats, pac, losses3);
    CKL();
    return 0;
}

int ptyb200_loss_grad(const ptyb200_cfg* c, const ptyb200_loss_cfg* lc, const float* dp, const float* meas_all,
                      const int64_t* idx, int32_t B, const double* stats, const float* pac, const float* upstream3,
                      float* G_out, const ptyb200_meas_cfg* mcfg, const float* meas_padded, ptyb200_stream s) {
    if (!c || !lc || !dp || !meas_all || !idx || !G_out) return fail_msg("NULL argument");
    cudaStream_t st = (cudaStream_t)s;
    LossK k = make_lossk(*lc);
    if (!stats || !upstream3) {
        // unscaled form for chunked steps (stats = upstream3 = NULL): the per-pixel factor of dL/dI only
        if (stats || upstream3) return fail_msg("loss_grad: stats and upstream3 must both be given, or both NULL (unscaled form)");
        if (k.b_on || (k.s_on != 0) == (k.p_on != 0))
            return fail_msg("the unscaled loss gradient (chunked steps) needs exactly one of loss_single / loss_poissn and no loss_pacbed");
        upstream3 = nullptr;
    }
    MeasView mv;
    if (const char* e = make_meas_view(*c, mcfg, meas_all, meas_padded, dp, G_out, &mv)) return fail_msg(e);
    unsigned chunks = (unsigned)((c->N * c->N + 256 * 8 - 1) / (256 * 8));
    k_loss_grad<<<dim3(chunks, B), 256, 0, st>>>(k, dp, mv, idx, B, c->N, stats, pac, upstream3, G_out);
    CKL();
    return 0;
}

int ptyb200_sparse_forward(const ptyb200_cfg* c, const ptyb200_loss_cfg* lc, const float* objp, const int32_t* crop_pos,
                           const int64_t* idx, int32_t B, const float* occu, float* loss_out, double* Ssum, int32_t* cover,
                           ptyb200_stream s) {
    if (!c || !lc || !objp || !crop_pos || !idx || !occu || !loss_out || !Ssum || !cover) return fail_msg("NULL argument");
    cudaStream_t st = (cudaStream_t)s;
    Dims d{c->N, c->P, c->M, c->Z, c->Noy, c->Nox, B, 0};
    CK(cudaMemsetAsync(Ssum, 0, sizeof(double) * c->M, st));
    CK(cudaMemsetAsync(cover, 0, (size_t)c->Noy * c->Nox * 4, st));
    k_cover<<<dim3((c->N * c->N + 1023) / 1024, B), 256, 0, st>>>(d, crop_pos, idx, cover);
    CKL();
    const size_t plane = (size_t)c->Noy * c->Nox;
    unsigned chunks = (unsigned)((plane + 256 * 8 - 1) / (256 * 8));
    k_sparse_partial<<<dim3(chunks, c->M * c->Z), 256, 0, st>>>(d, lc->sparse_order, objp, cover, Ssum);
    CKL();
    k_sparse_final<<<1, 32, 0, st>>>(d, lc->sparse_weight, lc->sparse_order, occu, Ssum, loss_out);
    CKL();
    return 0;
}

int ptyb200_sparse_grad(const ptyb200_cfg* c, const ptyb200_loss_cfg* lc, const float* objp, const int32_t* crop_pos,
                        const int64_t* idx, int32_t B, const float* occu, const double* Ssum, const float* upstream,
                        const int32_t* cover, float* g_objp, ptyb200_stream s) {
    if (!c || !lc || !objp || !crop_pos || !idx || !occu || !Ssum || !upstream || !cover || !g_objp) return fail_msg("NULL argument");
    cudaStream_t st = (cudaStream_t)s;
    Dims d{c->N, c->P, c->M, c->Z, c->Noy, c->Nox, B, 0};
    size_t n = (size_t)c->M * c->Z * c->Noy * c->Nox;
    k_sparse_grad<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, lc->sparse_weight, lc->sparse_order, objp, occu, Ssum, upstream, cover, g_objp);
    CKL();
    return 0;
}

namespace {
Blur5 make_blur5(float sigma) {
    Blur5 b;
    double sum = 0, k[5];
    for (int i = 0; i < 5; ++i) { const double t = (i - 2) / (double)sigma; k[i] = exp(-0.5 * t * t); sum += k[i]; }
    for (int i = 0; i < 5; ++i) b.k[i] = (float)(k[i] / sum);
    return b;
}
}  // namespace

int ptyb200_roi_blur(const ptyb200_cfg* c, const int64_t* idx, int32_t B, const float* obja, const float* objp, const int32_t* crop_pos,
                     float sigma, float* tmp, float* out_a, float* out_p, ptyb200_stream s) {
    if (int r = check_cfg(c, B)) return r;
    if (!idx || !obja || !objp || !crop_pos || !out_a || !out_p) return fail_msg("NULL argument");
    if (sigma < 0.f) return fail_msg("sigma must be >= 0");
    if (sigma > 0.f && !tmp) return fail_msg("roi_blur needs tmp (2 x B*M*Z*N*N floats) when sigma > 0");
    cudaStream_t st = (cudaStream_t)s;
    Dims d{c->N, c->P, c->M, c->Z, c->Noy, c->Nox, B, 0};
    const long long per = (long long)B * c->M * c->Z * c->N * c->N;
    const dim3 grid((unsigned)((per + 255) / 256), 2);
    if (sigma > 0.f) {
        const Blur5 b = make_blur5(sigma);
        k_roi_blurx<true><<<grid, 256, 0, st>>>(d, b, idx, crop_pos, obja, objp, tmp, tmp + per); CKL();
        k_blur5<1, false><<<grid.x, 256, 0, st>>>(b, tmp, out_a, per, c->N, c->N); CKL();
        k_blur5<1, false><<<grid.x, 256, 0, st>>>(b, tmp + per, out_p, per, c->N, c->N); CKL();
    } else {
        k_roi_blurx<false><<<grid, 256, 0, st>>>(d, Blur5{}, idx, crop_pos, obja, objp, out_a, out_p); CKL();
    }
    return 0;
}

int ptyb200_roi_blur_adjoint(const ptyb200_cfg* c, const int64_t* idx, int32_t B, const int32_t* crop_pos, float sigma, const float* g_out_a,
                             const float* g_out_p, float* tmp, float* g_obja, float* g_objp, ptyb200_stream s) {
    if (int r = check_cfg(c, B)) return r;
    if (!idx || !crop_pos) return fail_msg("NULL argument");
    if (sigma < 0.f) return fail_msg("sigma must be >= 0");
    if ((g_out_a && !g_obja) || (g_out_p && !g_objp)) return fail_msg("a patch gradient was given without its dense gradient buffer");
    if (sigma > 0.f && !tmp) return fail_msg("roi_blur_adjoint needs tmp (2 x B*M*Z*N*N floats) when sigma > 0");
    cudaStream_t st = (cudaStream_t)s;
    Dims d{c->N, c->P, c->M, c->Z, c->Noy, c->Nox, B, 0};
    const long long per = (long long)B * c->M * c->Z * c->N * c->N;
    const dim3 grid((unsigned)((per + 255) / 256), 2);
    if (sigma > 0.f) {
        const Blur5 b = make_blur5(sigma);
        if (g_out_a) { k_blur5<1, true><<<grid.x, 256, 0, st>>>(b, g_out_a, tmp, per, c->N, c->N); CKL(); }
        if (g_out_p) { k_blur5<1, true><<<grid.x, 256, 0, st>>>(b, g_out_p, tmp + per, per, c->N, c->N); CKL(); }
        k_roi_blurx_adj_scatter<true><<<grid, 256, 0, st>>>(d, b, idx, crop_pos, g_out_a ? tmp : nullptr, g_out_p ? tmp + per : nullptr, g_obja, g_objp); CKL();
    } else {
        k_roi_blurx_adj_scatter<false><<<grid, 256, 0, st>>>(d, Blur5{}, idx, crop_pos, g_out_a, g_out_p, g_obja, g_objp); CKL();
    }
    return 0;
}

int ptyb200_simlar_forward(const ptyb200_cfg* c, int32_t B, const float* plane, const float* occu, int32_t Zo, int32_t Yo, int32_t Xo,
                           float weight, double* sum_out, ptyb200_stream s) {
    if (int r = check_cfg(c, B)) return r;
    if (!plane || !occu || !sum_out) return fail_msg("NULL argument");
    if (c->M < 2 || c->M > 8) return fail_msg("loss_simlar needs 2 <= M <= 8 object modes");
    if (Zo < 1 || Yo < 1 || Xo < 1 || Zo > c->Z || Yo > c->N || Xo > c->N) return fail_msg("loss_simlar: pooled size must be in [1, input size]");
    SimlarDims d{B, c->M, c->Z, c->N, Zo, Yo, Xo};
    const long long cells = (long long)B * Zo * Yo * Xo;
    unsigned grid = (unsigned)((cells + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    k_simlar_fwd<<<grid, 256, 0, (cudaStream_t)s>>>(d, plane, occu, (double)weight / (double)cells, sum_out);
    CKL();
    return 0;
}

int ptyb200_simlar_backward(const ptyb200_cfg* c, int32_t B, const float* plane, const float* occu, int32_t Zo, int32_t Yo, int32_t Xo,
                            float weight, const float* upstream, float* g_plane, ptyb200_stream s) {
    if (int r = check_cfg(c, B)) return r;
    if (!plane || !occu || !upstream || !g_plane) return fail_msg("NULL argument");
    if (c->M < 2 || c->M > 8) return fail_msg("loss_simlar needs 2 <= M <= 8 object modes");
    if (Zo < 1 || Yo < 1 || Xo < 1 || Zo > c->Z || Yo > c->N || Xo > c->N) return fail_msg("loss_simlar: pooled size must be in [1, input size]");
    SimlarDims d{B, c->M, c->Z, c->N, Zo, Yo, Xo};
    const long long cells = (long long)B * Zo * Yo * Xo;
    unsigned grid = (unsigned)((cells + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    k_simlar_bwd<<<grid, 256, 0, (cudaStream_t)s>>>(d, plane, occu, (float)((double)weight / (double)cells), upstream, g_plane);
    CKL();
    return 0;
}

int ptyb200_gaussian_blur5(const float* in, float* tmp, float* out, int64_t planes, int32_t H, int32_t W, float sigma,
                           int32_t transpose, ptyb200_stream s) {
    if (!in || !tmp || !out) return fail_msg("NULL argument");
    if (planes < 1 || H < 3 || W < 3) return fail_msg("gaussian_blur5 needs planes >= 1 and H, W >= 3 (reflect padding of 2)");
    if (!(sigma > 0.f)) return fail_msg("sigma must be positive");
    if (in == out || in == tmp || tmp == out) return fail_msg("in, tmp and out must be distinct buffers");
    Blur5 b;
    double sum = 0, k[5];
    for (int i = 0; i < 5; ++i) { const double t = (i - 2) / (double)sigma; k[i] = exp(-0.5 * t * t); sum += k[i]; }
    for (int i = 0; i < 5; ++i) b.k[i] = (float)(k[i] / sum);
    const long long total = (long long)planes * H * W;
    const unsigned grid = (unsigned)((total + 255) / 256);
    cudaStream_t st = (cudaStream_t)s;
    // forward: x pass then y pass (as torchvision's separable kernel); the adjoint applies the transposed factors in reverse order
    if (!transpose) {
        k_blur5<0, false><<<grid, 256, 0, st>>>(b, in, tmp, total, H, W); CKL();
        k_blur5<1, false><<<grid, 256, 0, st>>>(b, tmp, out, total, H, W); CKL();
    } else {
        k_blur5<1, true><<<grid, 256, 0, st>>>(b, in, tmp, total, H, W); CKL();
        k_blur5<0, true><<<grid, 256, 0, st>>>(b, tmp, out, total, H, W); CKL();
    }
    return 0;
}

int ptyb200_loss_finalize(const ptyb200_cfg* c, const ptyb200_loss_cfg* lc, int32_t B_total, const double* stats, const float* pac,
                          float* losses3, ptyb200_stream s) {
    if (!c || !lc || !stats || !losses3 || B_total < 1) return fail_msg("NULL argument");
    LossK k = make_lossk(*lc);
    k_loss_final<<<1, 256, 0, (cudaStream_t)s>>>(k, B_total, c->N, const_cast<double*>(stats), const_cast<float*>(pac), losses3);
    CKL();
    return 0;
}

int ptyb200_loss_scale(const ptyb200_cfg* c, const ptyb200_loss_cfg* lc, int32_t B_total, const double* stats, const float* upstream3,
                       float* scale_out, ptyb200_stream s) {
    if (!c || !lc || !stats || !upstream3 || !scale_out || B_total < 1) return fail_msg("NULL argument");
    LossK k = make_lossk(*lc);
    if (k.b_on || (k.s_on != 0) == (k.p_on != 0))
        return fail_msg("loss_scale needs exactly one of loss_single / loss_poissn and no loss_pacbed");
    k_loss_scale<<<1, 32, 0, (cudaStream_t)s>>>(k, (double)B_total * c->N * c->N, stats, upstream3, scale_out);
    CKL();
    return 0;
}

int ptyb200_backward_zero(const ptyb200_cfg* c, int32_t B, void* workspace, float* g_probe, float* g_shifts, uint32_t need_mask,
                          ptyb200_stream s) {
    if (int r = check_cfg(c, B)) return r;
    if (!workspace) return fail_msg("NULL argument");
    if (c->reserved[1] & 1) return fail_msg("backward_zero does not cover patch mode");
    cudaStream_t st = (cudaStream_t)s;
    Workspace w = carve(*c, B, workspace);
    const size_t obj = (size_t)c->M * c->Z * c->Noy * c->Nox, pn = (size_t)c->P * c->N * c->N * 8;
    if (need_mask & PTYB200_NEED_OBJ) {
        if (use_fused(*c)) {
            fused128::Scratch sc = fused64::covers(*c) ? fused64::carve_scratch(*c, B, w.fused) : fused128::carve_scratch(*c, B, w.fused);
            CK(cudaMemsetAsync(sc.gOpack, 0, obj * 16, st));
        } else CK(cudaMemsetAsync(w.gO, 0, obj * 8, st));
    }
    if (need_mask & PTYB200_NEED_PROBE) {
        if (c->shift_probes) {
            CK(cudaMemsetAsync(w.gPhatT, 0, pn, st));
            if (use_fused(*c)) {
                fused128::Scratch sc = fused64::covers(*c) ? fused64::carve_scratch(*c, B, w.fused) : fused128::carve_scratch(*c, B, w.fused);
                CK(cudaMemsetAsync(sc.gPhatF, 0, pn, st));
            }
        } else {
            if (!g_probe) return fail_msg("g_probe is NULL");
            CK(cudaMemsetAsync(g_probe, 0, pn, st));
        }
    }
    if ((need_mask & PTYB200_NEED_SHIFTS) && g_shifts) CK(cudaMemsetAsync(g_shifts, 0, (size_t)c->Ntot * 2 * 4, st));
    return 0;
}

int ptyb200_accumulators_add(const ptyb200_cfg* c, int32_t B, void* workspace_dst, const void* workspace_src, uint32_t need_mask,
                             ptyb200_stream s) {
    if (int r = check_cfg(c, B)) return r;
    if (!workspace_dst || !workspace_src) return fail_msg("NULL argument");
    if (c->reserved[1] & 1) return fail_msg("accumulators_add does not cover patch mode");
    cudaStream_t st = (cudaStream_t)s;
    Workspace wd = carve(*c, B, workspace_dst), ws = carve(*c, B, const_cast<void*>(workspace_src));
    const size_t obj = (size_t)c->M * c->Z * c->Noy * c->Nox, pn8 = (size_t)c->P * c->N * c->N * 8;
    auto add = [&](void* d, const void* a, size_t bytes) -> int {
        const size_t n4 = bytes / 16;                                  // every accumulator is a multiple of 16 bytes, 256-byte aligned
        k_add_into<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>((float4*)d, (const float4*)a, n4);
        CKL();
        return 0;
    };
    if (need_mask & PTYB200_NEED_OBJ) {
        if (use_fused(*c)) {
            fused128::Scratch a = fused64::covers(*c) ? fused64::carve_scratch(*c, B, wd.fused) : fused128::carve_scratch(*c, B, wd.fused);
            fused128::Scratch b = fused64::covers(*c) ? fused64::carve_scratch(*c, B, ws.fused) : fused128::carve_scratch(*c, B, ws.fused);
            if (int r = add(a.gOpack, b.gOpack, obj * 16)) return r;
        } else if (int r = add(wd.gO, ws.gO, (obj * 8 + 15) & ~size_t(15))) return r;
    }
    if ((need_mask & PTYB200_NEED_PROBE) && c->shift_probes) {
        if (use_fused(*c)) {
            fused128::Scratch a = fused64::covers(*c) ? fused64::carve_scratch(*c, B, wd.fused) : fused128::carve_scratch(*c, B, wd.fused);
            fused128::Scratch b = fused64::covers(*c) ? fused64::carve_scratch(*c, B, ws.fused) : fused128::carve_scratch(*c, B, ws.fused);
            if (int r = add(a.gPhatF, b.gPhatF, pn8)) return r;
        } else if (int r = add(wd.gPhatT, ws.gPhatT, pn8)) return r;
    }
    return 0;
}

int ptyb200_backward_finish(const ptyb200_cfg* c, int32_t B, const float* obja, const float* objp, void* workspace, float* g_obja,
                            float* g_objp, float* g_probe, float* g_shifts, uint32_t need_mask, const float* scale, ptyb200_stream s) {
    if (int r = check_cfg(c, B)) return r;
    if (!obja || !objp || !workspace) return fail_msg("NULL argument");
    if (c->reserved[1] & 1) return fail_msg("chunked steps do not cover patch mode");
    if (need_mask & (PTYB200_NEED_TILTS | PTYB200_NEED_DZ)) return fail_msg("chunked steps do not cover tilt / thickness gradients");
    cudaStream_t st = (cudaStream_t)s;
    Workspace w = carve(*c, B, workspace);
    const bool need_obj = (need_mask & PTYB200_NEED_OBJ) != 0, need_probe = (need_mask & PTYB200_NEED_PROBE) != 0;
    const bool need_shift = (need_mask & PTYB200_NEED_SHIFTS) && c->shift_probes;
    if (need_obj && (!g_obja || !g_objp)) return fail_msg("g_obja/g_objp is NULL");
    if (need_probe && !g_probe) return fail_msg("g_probe is NULL");
    const size_t obj = (size_t)c->M * c->Z * c->Noy * c->Nox, pn = (size_t)c->P * c->N * c->N * 2;
    DISPATCH_N(c->N, {
        if (use_fused(*c)) {
            BwdArgs a;
            memset(&a, 0, sizeof a);
            a.f = make_fwd_args(*c, B, w, nullptr, nullptr, nullptr, nullptr, nullptr);
            a.need_obj = need_obj; a.need_probe = need_probe;
            if (int r = fused64::covers(*c)
                            ? fused64::backward(*c, B, a, obja, objp, g_obja, g_objp, w.fused, (float2*)g_probe, w.gPhatT, st, g_err, &g_launches, c->reserved[4] & PTYB200_ACC_ADD_OBJ, scale, true)
                            : fused128::backward(*c, B, a, obja, objp, g_obja, g_objp, w.fused, (float2*)g_probe, w.gPhatT, st, g_err, &g_launches, c->reserved[4] & PTYB200_ACC_ADD_OBJ, scale, true)) return r;
        } else if (need_obj) {
            k_obj_finish<<<(unsigned)((obj + 255) / 256), 256, 0, st>>>(w.gO, obja, objp, g_obja, g_objp, obj, scale, (c->reserved[4] & PTYB200_ACC_ADD_OBJ) ? 1 : 0);
            CKL();
        }
        if (need_probe && c->shift_probes)
            if (int r = fft2_tiles<F>(w.gPhatT, w.tmpP, (float2*)g_probe, c->P, +1, st)) return r;
    });
    if (scale) {
        if (need_probe) { k_scale<<<(unsigned)((pn + 255) / 256), 256, 0, st>>>(g_probe, pn, scale); CKL(); }
        if (need_shift && g_shifts) { const size_t n = (size_t)c->Ntot * 2; k_scale<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g_shifts, n, scale); CKL(); }
    }
    return 0;
}

int ptyb200_sparse_groups(const double* pos_ordered, int32_t n, int32_t G, int32_t* labels_out, ptyb200_stream s) {
    if (!pos_ordered || !labels_out) return fail_msg("NULL argument");
    if (G < 1 || n < G) return fail_msg("sparse_groups needs 1 <= G <= n (one seed per group comes first)");
    if (G > 4096) return fail_msg("sparse_groups: at most 4096 groups");
    const size_t smem = (size_t)G * 8 + 32 * 8 + 32 * 4;
    k_sparse_groups<<<1, GROUP_THREADS, smem, (cudaStream_t)s>>>(reinterpret_cast<const double2*>(pos_ordered), n, G, labels_out);
    CKL();
    return 0;
}

int ptyb200_blur_axis(const float* in, float* out, int64_t outer, int32_t L, int64_t inner, int32_t ksize, float sigma,
                      int32_t pad_mode, ptyb200_stream s) {
    if (!in || !out || in == out) return fail_msg("blur_axis needs distinct in / out buffers");
    if (outer < 1 || L < 1 || inner < 1) return fail_msg("blur_axis: empty array");
    if (ksize < 1 || ksize > 15 || !(ksize & 1)) return fail_msg("blur_axis: kernel size must be odd and <= 15");
    if (!(sigma > 0.f)) return fail_msg("sigma must be positive");
    if (pad_mode == 0 && ksize / 2 > L - 1) return fail_msg("blur_axis: reflect padding needs kernel_size/2 < length");
    BlurTaps t;
    double k[15], sum = 0;
    for (int i = 0; i < ksize; ++i) { const double x = (i - (ksize - 1) / 2.0) / (double)sigma; k[i] = exp(-0.5 * x * x); sum += k[i]; }
    for (int i = 0; i < ksize; ++i) t.k[i] = (float)(k[i] / sum);
    t.n = ksize;
    const long long total = (long long)outer * L * inner;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (pad_mode == 0) k_blur_axis<0><<<grid, 256, 0, (cudaStream_t)s>>>(t, in, out, total, L, inner);
    else k_blur_axis<1><<<grid, 256, 0, (cudaStream_t)s>>>(t, in, out, total, L, inner);
    CKL();
    return 0;
}

int ptyb200_object_constraints(const ptyb200_obj_constraints* oc, float* obja, float* objp, int64_t n, float* scratch, ptyb200_stream s) {
    if (!oc || !obja || !objp || n < 1) return fail_msg("NULL argument");
    cudaStream_t st = (cudaStream_t)s;
    ObjConstraints c;
    c.mirrored_on = oc->mirrored_on; c.mirrored_relax = oc->mirrored_relax; c.mirrored_scale = oc->mirrored_scale; c.mirrored_power = oc->mirrored_power;
    c.thresh_on = oc->thresh_on; c.thresh_relax = oc->thresh_relax; c.thresh_lo = oc->thresh_lo; c.thresh_hi = oc->thresh_hi;
    c.postiv_on = oc->postiv_on; c.postiv_relax = oc->postiv_relax; c.postiv_subtract_min = oc->postiv_subtract_min;
    if (!c.mirrored_on && !c.thresh_on && !c.postiv_on) return 0;
    if (c.postiv_on && c.postiv_subtract_min) {
        if (!scratch) return fail_msg("objp_postiv mode 'subtract_min' needs one float of scratch");
        const unsigned inf_bits = 0x7f800000u;
        CK(cudaMemsetAsync(scratch, 0, 4, st));
        CK(cudaMemcpyAsync(scratch, &inf_bits, 4, cudaMemcpyHostToDevice, st));
        k_obj_min<<<148 * 4, 256, 0, st>>>(objp, n, scratch);
        CKL();
    }
    k_obj_voxel_constraints<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c, obja, objp, n, scratch);
    CKL();
    return 0;
}

int ptyb200_adam_step(int32_t count, float* const* params, const float* const* grads, float* const* exp_avg,
                      float* const* exp_avg_sq, float* const* steps, const float* lrs, const int64_t* numels, float beta1, float beta2,
                      float eps, ptyb200_stream s) {
    if (count < 1 || count > 8) return fail_msg("adam_step handles 1..8 tensors per call");
    if (!params || !grads || !exp_avg || !exp_avg_sq || !steps || !lrs || !numels) return fail_msg("NULL argument");
    AdamTensors a;
    long long nmax = 0;
    for (int i = 0; i < count; ++i) {
        a.p[i] = params[i]; a.g[i] = grads[i]; a.m[i] = exp_avg[i]; a.v[i] = exp_avg_sq[i];
        a.step[i] = steps[i]; a.lr[i] = lrs[i]; a.n[i] = numels[i];
        if (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i] || !steps[i]) return fail_msg("NULL tensor pointer");
        if (numels[i] > nmax) nmax = numels[i];
    }
    a.count = count; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    cudaStream_t st = (cudaStream_t)s;
    k_adam_advance<<<1, 32, 0, st>>>(a);
    CKL();
    unsigned bx = (unsigned)((nmax + 1023) / 1024);
    if (bx > 148 * 8) bx = 148 * 8;
    if (bx < 1) bx = 1;
    k_adam<<<dim3(bx, count), 256, 0, st>>>(a);
    CKL();
    return 0;
}

}  // extern "C"
