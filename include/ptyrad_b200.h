/* ptyrad_b200 -- C ABI of the B200-native multislice ptychography hot path.
 *
 * The reference (wdwzyyg/ptyrad, pure Python/PyTorch) has no FFI of its own: its hot path is the Python
 * surface PtychoAD.forward -> multislice_forward_model_vec_all -> CombinedLoss.forward -> autograd.
 * Each entry point below names the reference function(s) it replaces (file:line relative to the reference
 * tree).  All pointers are DEVICE pointers owned by the caller (torch tensors in practice); nothing is
 * allocated inside the library; work is enqueued on the caller's stream and the calls do not block.
 * Return value: 0 = OK, non-zero = error, text via ptyb200_last_error().
 *
 * State the library keeps between calls (process-wide, created on first use): the optional timing events (ptyb200_timing_*), the
 * launch counter, and two internal non-blocking streams + events on which ptyb200_forward* / ptyb200_backward run their short,
 * mutually independent setup and completion kernels as parallel branches.  The branches are forked from and joined back into the
 * caller's stream inside the same call, so the caller still sees ONE stream and a stream capture sees a fork / join subgraph;
 * PTYB200_NO_BRANCHES=1 in the environment keeps everything on the caller's stream, and a call made on another device than the one
 * the streams were created on does the same.
 *
 * Complex arrays are interleaved (re,im) float32 pairs ("float2").
 */
#ifndef PTYRAD_B200_H
#define PTYRAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTYB200_ABI_VERSION 3

/* cudaStream_t without pulling in the CUDA headers */
typedef void* ptyb200_stream;

/* bits of `need_mask` (which gradients the caller wants; mirrors requires_grad, reconstruction.py:783-790) */
#define PTYB200_NEED_OBJ     1u   /* obja + objp */
#define PTYB200_NEED_PROBE   2u
#define PTYB200_NEED_SHIFTS  4u
#define PTYB200_NEED_TILTS   8u
#define PTYB200_NEED_DZ     16u

/* code paths (ptyb200_cfg.path) */
#define PTYB200_PATH_AUTO    0    /* fused on-chip kernels where available (N == 128, N == 64), else general */
#define PTYB200_PATH_GENERAL 1    /* row/column-pass kernels, any supported N */
#define PTYB200_PATH_FUSED   2    /* fail if the fused kernels do not cover this configuration */

typedef struct ptyb200_cfg {
    int32_t N;             /* pattern = probe = ROI size in px (square, even; supported: 16,32,48,64,96,128,192,256) */
    int32_t P;             /* probe modes   */
    int32_t M;             /* object modes  */
    int32_t Z;             /* slices        */
    int32_t Noy, Nox;      /* object canvas */
    int32_t Ntot;          /* scan positions (rows of crop_pos / shifts / measurements) */
    int32_t shift_probes;  /* 1: sub-pixel Fourier shift of the probe per position (models.py:120,294-295) */
    int32_t tilt_mode;     /* 0: propagator shared by all positions; 1: one global tilt (1,2); 2: per-position (Ntot,2) */
    int32_t stash_fourier; /* 1: keep the Fourier-domain waves so tilt / thickness gradients can be formed */
    int32_t path;          /* PTYB200_PATH_* */
    int32_t reserved[5];   /* [0] workspace batch capacity: the B the workspace was sized for (ptyb200_workspace_bytes) when calls
                                  run on chunks of <= that many samples that share one workspace (chunked steps), 0 = the call's B;
                              [1] bit 0: PATCH MODE -- obja/objp (and their gradients) are per-sample ROI stacks (B,M,Z,N,N),
                                  e.g. pre-blurred patches (models.py:275-284); needs Noy == Nox == N, crop_pos is ignored;
                              [2] general path: samples per chunk (the slice sequence runs chunk by chunk so that the pass buffers
                                  stay L2-resident), 0 = heuristic;  [3] general path: probe modes per CTA, 0 = heuristic.
                                  [2] and [3] change the workspace size: use the same cfg for ptyb200_workspace_bytes;
                              [4] PTYB200_ACC_* flags of a chunked step (below), 0 = a whole batch per call */
    float   dx;            /* real-space pixel size (propagator k-grid, models.py:164-171) */
    float   lambd;         /* wavelength (Kz, models.py:222-223) */
    float   eps;           /* added to the intensities after the mode sum (forward.py:79); reference: 1e-10 */
    float   reserved_f;
} ptyb200_cfg;

/* Chunked steps: a batch larger than the workspace allows is processed chunk by chunk (forward, unscaled loss gradient, adjoint per
 * chunk; only one chunk's wave stash is alive), the way SURVEY 8e asks ("memory must not grow with B").  The data losses are not
 * additive over chunks -- loss_single = sqrt(mean((I^p - M^p)^2)) / mean(M^p) over the WHOLE batch (losses.py:42-47) -- but dL/dI
 * factors into (a scalar of the batch sums) x (a per-pixel term), and the adjoint is linear in dL/dI: every chunk runs its adjoint on
 * the per-pixel term (ptyb200_loss_grad with stats = upstream = NULL) into shared accumulators, and ptyb200_backward_finish applies
 * the scalar (ptyb200_loss_scale, from the completed sums) when it forms the gradients.  cfg.reserved[4] carries: */
#define PTYB200_ACC_KEEP_STATS 1   /* ptyb200_forward_loss: add to the loss sums of the earlier chunks (no zeroing, no final scalars) */
#define PTYB200_ACC_KEEP_GRADS 2   /* ptyb200_backward: add to the gradient accumulators of the earlier chunks (no zeroing) */
#define PTYB200_ACC_NO_FINISH  4   /* ptyb200_backward: leave the accumulators raw; ptyb200_backward_finish completes them */
/* Two more bits let a caller take short, independent launches off the critical path of a step (run them on a second stream while the
 * multislice forward occupies the first): */
#define PTYB200_ACC_NO_LOSS_FINAL 8 /* ptyb200_forward_loss: leave losses3 to a later ptyb200_loss_finalize */
#define PTYB200_ACC_ADD_OBJ   16   /* ptyb200_backward(_finish): ADD the object gradients to g_obja / g_objp, which the caller zeroed and
                                      may have pre-loaded with other terms (ptyb200_sparse_grad), instead of overwriting them */

typedef struct ptyb200_loss_cfg {
    /* CombinedLoss terms computed natively (losses.py:36-104); state 0 => term is 0 */
    int32_t single_state;  float single_weight;  float single_pow;                   /* losses.py:36-50  */
    int32_t poissn_state;  float poissn_weight;  float poissn_pow;  float poissn_eps; /* losses.py:52-75  */
    int32_t pacbed_state;  float pacbed_weight;  float pacbed_pow;                   /* losses.py:77-89  */
    int32_t sparse_state;  float sparse_weight;  float sparse_order;                 /* losses.py:91-104 */
} ptyb200_loss_cfg;

/* How the loss sees the measured patterns (PtychoAD.get_measurements, models.py:384-416).  NULL = rows of the stored (Ntot,N,N)
 * array as they are.  Otherwise: rows are (Hs,Ws); if Hp/Wp > 0 each row is pasted into the background canvas `meas_padded`
 * (Hp,Wp) at [h1:h2, w1:w2] ("on-the-fly" padding, models.py:401-405); if a scale factor differs from 1 the result is resampled
 * bilinearly (torch interpolate(mode='bilinear', scale_factor=...), align_corners = False) and divided by scale_y*scale_x
 * (models.py:407-409).  The final size must equal cfg.N.  Evaluated inside the loss kernels; nothing is materialised. */
typedef struct ptyb200_meas_cfg {
    int32_t Hs, Ws;          /* stored pattern size */
    int32_t Hp, Wp;          /* padded canvas size, 0 = no padding */
    int32_t h1, h2, w1, w2;  /* paste window */
    float   scale_y, scale_x;/* resample factors, 0 or 1 = none */
} ptyb200_meas_cfg;

int         ptyb200_abi_version(void);
const char* ptyb200_last_error(void);

/* Instrumentation for bench.py (no reference counterpart): number of kernels this library has launched so far, and
 * CUDA-event timing of the multislice forward / adjoint sections on the caller's stream (the sections exclude the
 * small per-step setup and finish kernels).  timing_read synchronises on the recorded events, returns the summed
 * milliseconds and the number of sections since the last read (at most 256 are kept), and resets. */
long long   ptyb200_launch_count(void);
void        ptyb200_timing_enable(int on);
int         ptyb200_timing_read(double* ms_forward, double* ms_backward, int* n_forward, int* n_backward);

/* Bytes of caller-owned scratch one forward/backward pair of batch size B needs (wave stash for the
 * adjoint, transposed wave buffers, complex object, gradient scratch).  One workspace per in-flight
 * forward: the LBFGS closure (reconstruction.py:705-718) keeps several alive. */
size_t ptyb200_workspace_bytes(const ptyb200_cfg* cfg, int32_t B);

/* H = exp(i*dz*Kz) on the half-bin-shifted grid, evaluated in float64 and rounded once
 * (replaces the float32 torch.exp(1j*dz*Kz) of models.py:341,355; Kz from models.py:222-223). */
int ptyb200_propagator(const ptyb200_cfg* cfg, const float* dz, float* H_out /* (N,N) float2 */, ptyb200_stream s);

/* Bit-exact ROI gather (models.py:251-265): patches[b,m,z,y,x,{a,phi}] = obj{a,p}[m,z,crop[idx[b]].y+y, crop[idx[b]].x+x] */
int ptyb200_gather_patches(const ptyb200_cfg* cfg, const int64_t* idx, int32_t B, const float* obja, const float* objp,
                           const int32_t* crop_pos, float* patches_out, ptyb200_stream s);

/* PtychoAD.forward (models.py:422-436) = get_obj_ROI + get_probes (image_proc.py:495-537) + get_propagators
 * (models.py:300-360) + multislice_forward_model_vec_all (forward.py:20-80).
 *   idx      (B)            int64   scan indices of the batch
 *   obja,objp(M,Z,Noy,Nox)  float32 amplitude / phase
 *   crop_pos (Ntot,2)       int32   ROI top-left (y,x)
 *   probe    (P,N,N)        float2
 *   shifts   (Ntot,2)       float32 sub-pixel (y,x) shifts; may be NULL iff !shift_probes
 *   Hbase    (N,N)          float2  propagator for zero tilt (model.H, or ptyb200_propagator output)
 *   tilts    (1|Ntot,2)     float32 mrad (y,x); NULL iff tilt_mode == 0
 *   dz       scalar         float32 slice thickness (device; used by the tilt ramp only)
 *   occu     (M)            float32
 *   dp_out   (B,N,N)        float32 fftshifted intensities + eps
 */
int ptyb200_forward(const ptyb200_cfg* cfg, const int64_t* idx, int32_t B, const float* obja, const float* objp,
                    const int32_t* crop_pos, const float* probe, const float* shifts, const float* Hbase,
                    const float* tilts, const float* dz, const float* occu, float* dp_out, void* workspace,
                    ptyb200_stream s);

/* ptyb200_forward with the mode reduction FUSED with the data losses (forward.py:79 + losses.py:36-89; north star item 3): the kernel
 * that completes a pattern's intensities also adds the pattern's contribution to the batch sums, so dp is not re-read by a reduction
 * launch; losses3 / stats / pacbed_scratch as in ptyb200_loss_forward (ptyb200_loss_grad then takes the same stats).  The measured
 * pattern of sample b is row meas_rows[b] of meas_all (NULL: idx[b]; a rank that holds only its shard passes its own row map). */
int ptyb200_forward_loss(const ptyb200_cfg* cfg, const int64_t* idx, int32_t B, const float* obja, const float* objp,
                         const int32_t* crop_pos, const float* probe, const float* shifts, const float* Hbase,
                         const float* tilts, const float* dz, const float* occu, float* dp_out, void* workspace,
                         const ptyb200_loss_cfg* lc, const float* meas_all, const int64_t* meas_rows, const ptyb200_meas_cfg* mcfg,
                         const float* meas_padded, float* losses3, double* stats, float* pacbed_scratch, ptyb200_stream s);

/* Adjoint of ptyb200_forward (what torch autograd derives for the reference).  Must follow
 * a forward on the same workspace with the same inputs.  G = dL/d(dp) (B,N,N).  Gradient outputs are DENSE
 * and are OVERWRITTEN (zero-filled where no pattern contributes), as the reference's .grad tensors are:
 *   g_obja,g_objp (M,Z,Noy,Nox); g_probe (P,N,N) float2; g_shifts (Ntot,2); g_tilts (1|Ntot,2); g_dz scalar.
 * Outputs whose bit is absent from need_mask may be NULL and cost nothing.  Alignment: g_obja, g_objp, g_probe 16 bytes,
 * g_shifts, g_tilts 8 bytes (vector stores / reductions); a misaligned pointer is refused with an error. */
int ptyb200_backward(const ptyb200_cfg* cfg, const int64_t* idx, int32_t B, const float* obja, const float* objp,
                     const int32_t* crop_pos, const float* probe, const float* shifts, const float* Hbase,
                     const float* tilts, const float* dz, const float* occu, const float* G, void* workspace,
                     float* g_obja, float* g_objp, float* g_probe, float* g_shifts, float* g_tilts, float* g_dz,
                     uint32_t need_mask, ptyb200_stream s);

/* CombinedLoss data terms (losses.py:36-89) on the batch: losses3 = {single, poissn, pacbed} (device floats).
 * `stats` is 8 doubles of caller-owned device memory kept for ptyb200_loss_grad.  meas_all is the full
 * (Ntot,N,N) measurement array; rows idx[b] are read in place (no gathered copy; models.py:399). */
int ptyb200_loss_forward(const ptyb200_cfg* cfg, const ptyb200_loss_cfg* lc, const float* dp, const float* meas_all,
                         const int64_t* idx, int32_t B, float* losses3, double* stats, float* pacbed_scratch /* 2*N*N */,
                         const ptyb200_meas_cfg* mcfg /* NULL: plain */, const float* meas_padded /* (Hp,Wp) or NULL */,
                         ptyb200_stream s);

/* PtychoAD.get_measurements(indices) as a tensor (models.py:384-416): out (B,N,N) = the padded / resampled rows idx. */
int ptyb200_gather_measurements(const ptyb200_cfg* cfg, const ptyb200_meas_cfg* mcfg, const float* meas_all, const float* meas_padded,
                                const int64_t* idx, int32_t B, float* out, ptyb200_stream s);

/* G = sum_t upstream[t] * d(loss_t)/d(dp), t over {single, poissn, pacbed}; upstream = 3 device floats. */
int ptyb200_loss_grad(const ptyb200_cfg* cfg, const ptyb200_loss_cfg* lc, const float* dp, const float* meas_all,
                      const int64_t* idx, int32_t B, const double* stats, const float* pacbed_scratch,
                      const float* upstream3, float* G_out, const ptyb200_meas_cfg* mcfg, const float* meas_padded,
                      ptyb200_stream s);

/* loss_sparse (losses.py:91-104) evaluated on the ROIs without materialising the patches:
 * loss = weight * sum_m occu_m * (mean_{b,z,y,x} |phi_patch|^n)^(1/n);  Ssum (M doubles) kept for the gradient. */
int ptyb200_sparse_forward(const ptyb200_cfg* cfg, const ptyb200_loss_cfg* lc, const float* objp, const int32_t* crop_pos,
                           const int64_t* idx, int32_t B, const float* occu, float* loss_out, double* Ssum,
                           int32_t* cover /* Noy*Nox int32, filled: ROIs of the batch covering each pixel */, ptyb200_stream s);

/* g_objp (dense, += ) gets upstream * d(loss_sparse)/d(objp); cover = the map ptyb200_sparse_forward filled. */
int ptyb200_sparse_grad(const ptyb200_cfg* cfg, const ptyb200_loss_cfg* lc, const float* objp, const int32_t* crop_pos,
                        const int64_t* idx, int32_t B, const float* occu, const double* Ssum, const float* upstream,
                        const int32_t* cover, float* g_objp, ptyb200_stream s);

/* 5x5 Gaussian blur with reflect padding over the last two dims of `planes` (H,W) float32 planes: what torchvision's
 * gaussian_blur(kernel_size=5, sigma) computes for the object pre-blur (models.py:275-284), the detector blur (models.py:379-380)
 * and loss_simlar (losses.py:125,134).  transpose = 1 applies the adjoint operator (the backward pass).  `tmp` is caller-owned
 * scratch of the same size; in, tmp and out must be distinct. */
int ptyb200_gaussian_blur5(const float* in, float* tmp, float* out, int64_t planes, int32_t H, int32_t W, float sigma,
                           int32_t transpose, ptyb200_stream s);

/* Object pre-blur without the gather tensor (models.py:251-284): out_a / out_p (B,M,Z,N,N) = the ROI planes of amplitude / phase for the
 * batch, 5x5 Gaussian-blurred with reflect padding inside the patch when sigma > 0 (sigma = 0: the plain ROI planes, what
 * get_obj_ROI gathers).  tmp = 2 * B*M*Z*N*N floats of scratch (sigma > 0).  These planes are what the multislice kernels take as their
 * object in patch mode (cfg.reserved[1]) and what loss_simlar works on.
 * ptyb200_roi_blur_adjoint: the backward pass -- patch gradients (either may be NULL) through the adjoint blur, scatter-ADDED into the
 * dense (M,Z,Noy,Nox) gradients (the caller zeroes them; atomics, like the reference's index_put_(accumulate=True)). */
int ptyb200_roi_blur(const ptyb200_cfg* cfg, const int64_t* idx, int32_t B, const float* obja, const float* objp, const int32_t* crop_pos,
                     float sigma, float* tmp, float* out_a, float* out_p, ptyb200_stream s);
int ptyb200_roi_blur_adjoint(const ptyb200_cfg* cfg, const int64_t* idx, int32_t B, const int32_t* crop_pos, float sigma,
                             const float* g_out_a, const float* g_out_p, float* tmp, float* g_obja, float* g_objp, ptyb200_stream s);

/* loss_simlar (losses.py:106-141) on one set of ROI planes (B,M,Z,N,N) from ptyb200_roi_blur (sigma = blur_std or 0): area
 * interpolation (adaptive average pooling to (Zo,Yo,Xo) = floor(size * scale_factor)), times occu_m, unbiased std over the M object
 * modes, mean over the pooled batch volume.  forward: sum_out[0] (double) += weight * that mean.  backward: g_plane (B,M,Z,N,N), zeroed
 * by the caller, += upstream[0] * d/d(plane); ptyb200_roi_blur_adjoint takes it to the dense gradients.  2 <= M <= 8. */
int ptyb200_simlar_forward(const ptyb200_cfg* cfg, int32_t B, const float* plane, const float* occu, int32_t Zo, int32_t Yo, int32_t Xo,
                           float weight, double* sum_out, ptyb200_stream s);
int ptyb200_simlar_backward(const ptyb200_cfg* cfg, int32_t B, const float* plane, const float* occu, int32_t Zo, int32_t Yo, int32_t Xo,
                            float weight, const float* upstream, float* g_plane, ptyb200_stream s);

/* Per-iteration object constraints, in place, no .data re-binding:
 * ptyb200_blur_axis: 1-D Gaussian (odd kernel_size <= 15, weights exp(-x^2/2 sigma^2)/sum) along the middle axis of an array viewed as
 *   (outer, L, inner); pad_mode 0 = reflect (obj_rblur: x pass then y pass = torchvision gaussian_blur, constraints.py:83-99),
 *   1 = replicate (obj_zblur along z: gaussian_blur_1d, constraints.py:101-114, utils/image_proc.py:443-455).  in != out.
 * ptyb200_object_constraints: mirrored_amp -> obja_thresh -> objp_postiv (constraints.py:165-208) in ONE pass over the n voxels of
 *   obja / objp; `scratch` = one device float (only for objp_postiv mode 'subtract_min'). */
typedef struct ptyb200_obj_constraints {
    int32_t mirrored_on; float mirrored_relax, mirrored_scale, mirrored_power;   /* constraints.py:165-178 */
    int32_t thresh_on;   float thresh_relax, thresh_lo, thresh_hi;               /* constraints.py:180-189 */
    int32_t postiv_on;   float postiv_relax; int32_t postiv_subtract_min;        /* constraints.py:191-208 */
} ptyb200_obj_constraints;
int ptyb200_blur_axis(const float* in, float* out, int64_t outer, int32_t L, int64_t inner, int32_t kernel_size, float sigma,
                      int32_t pad_mode, ptyb200_stream s);
int ptyb200_object_constraints(const ptyb200_obj_constraints* oc, float* obja, float* objp, int64_t n, float* scratch,
                               ptyb200_stream s);

/* Zeroes what a following ptyb200_backward(cfg.reserved[4] & PTYB200_ACC_KEEP_GRADS) accumulates into (the workspace accumulators, the
 * dense shift gradient, the probe gradient of unshifted probes): the same memsets ptyb200_backward starts with, as a call of their
 * own so that they can run early on another stream. */
int ptyb200_backward_zero(const ptyb200_cfg* cfg, int32_t B, void* workspace, float* g_probe, float* g_shifts, uint32_t need_mask,
                          ptyb200_stream s);

/* workspace_dst's gradient accumulators += workspace_src's (two parts of one batch that ran their adjoints concurrently on two
 * streams, each into its own workspace; both sized with the same cfg and B). */
int ptyb200_accumulators_add(const ptyb200_cfg* cfg, int32_t B, void* workspace_dst, const void* workspace_src, uint32_t need_mask,
                             ptyb200_stream s);

/* Completion of a chunked step (see PTYB200_ACC_*): object polar backward, probe-spectrum inverse FFT and the batch-level `scale`
 * (device float, e.g. from ptyb200_loss_scale; NULL = 1) applied to every requested gradient.  Same buffers / need_mask as the
 * ptyb200_backward calls whose accumulators (in `workspace`) it completes; tilt / thickness gradients are not available chunked. */
int ptyb200_backward_finish(const ptyb200_cfg* cfg, int32_t B, const float* obja, const float* objp, void* workspace, float* g_obja,
                            float* g_objp, float* g_probe, float* g_shifts, uint32_t need_mask, const float* scale, ptyb200_stream s);
/* losses3 = the data-loss scalars from the sums in `stats` accumulated over B_total samples (what ptyb200_forward_loss does at the
 * end of an unchunked call). */
int ptyb200_loss_finalize(const ptyb200_cfg* cfg, const ptyb200_loss_cfg* lc, int32_t B_total, const double* stats, const float* pac,
                          float* losses3, ptyb200_stream s);
/* scale_out[0] = upstream * (the factor of dL/dI that depends on the sums over all B_total samples); exactly one of loss_single /
 * loss_poissn must be active and loss_pacbed off (its gradient is not separable per pattern). */
int ptyb200_loss_scale(const ptyb200_cfg* cfg, const ptyb200_loss_cfg* lc, int32_t B_total, const double* stats, const float* upstream3,
                       float* scale_out, ptyb200_stream s);

/* 'sparse' grouping of scan positions (make_batches, reconstruction.py:540-587): the greedy assignment loop.  pos_ordered = (n,2)
 * float64 positions, the G group seeds first (group g's seed in row g: the point closest to the centroid of compact group g,
 * reconstruction.py:556-562), then the remaining points in the order the reference visits them.  labels_out[i] = group of row i:
 * each point joins the group whose nearest member is farthest from it (first group on ties, np.argmax).  One CTA; O(n^2 / 2). */
int ptyb200_sparse_groups(const double* pos_ordered, int32_t n, int32_t G, int32_t* labels_out, ptyb200_stream s);

/* optimizer.step() for torch.optim.Adam defaults (reconstruction.py:759; built at reconstruction.py:285-368):
 * one launch over up to 8 tensors with per-tensor learning rates.  The host arrays of pointers / lrs / numels are read
 * during the call.  steps[i] is tensor i's OWN step counter, one device float32 scalar (torch.optim.Adam's state['step']): the call
 * advances each by one before use (per-tensor bias correction, as torch does for tensors that join late through start_iter,
 * reconstruction.py:783-790), so no host synchronisation is needed. */
int ptyb200_adam_step(int32_t count, float* const* params, const float* const* grads, float* const* exp_avg,
                      float* const* exp_avg_sq, float* const* steps, const float* lrs, const int64_t* numels, float beta1,
                      float beta2, float eps, ptyb200_stream s);

#ifdef __cplusplus
}
#endif
#endif /* PTYRAD_B200_H */
